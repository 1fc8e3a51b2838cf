"""
Batch sources for fusionnet_main.train / radarnet_main.train and the device side of the reference's data path
(SURVEY.md 8f row 4).

The reference decodes every PNG to float32 on CPU workers, crops, and ships 4-byte values to the GPU
(src/data_utils.py:167-198, 238-318; src/datasets.py:19-109, 399-440).  Here the host only does what must be done on
the host -- file read + PNG inflate (Pillow) and the crop DECISION -- and keeps the on-disk sample types (uint8 RGB,
uint16 maps: 11 bytes per pixel of a 5-tensor sample instead of 28); the value codec (/ multiplier, <= 0 -> 0), the
HWC -> CHW layout change and, optionally, the crop itself run batched on the device (rcfd_decode_crop).  Nothing here
imports the reference.  ``'synthetic'`` paths produce the seeded synthetic batches of SURVEY 8d.
"""
import numpy as np
import torch

from . import ops, synth

DEPTH_MULTIPLIER = 256.0          # src/data_utils.py:238 (load_depth; FusionNetTrainingDataset also reads the response with it, src/datasets.py:411-414)
RESPONSE_MULTIPLIER = 2.0 ** 14   # src/data_utils.py:288 (load_response / save_response)


def read_paths(filepath):
    """Newline-delimited list of paths, up to the first empty line (reference src/data_utils.py:128-150)."""
    paths = []
    with open(filepath) as f:
        for line in f:
            line = line.rstrip('\n')
            if line == '':
                break
            paths.append(line)
    return paths


def load_image_u8(path):
    """H x W x 3 uint8, the samples behind the reference's load_image (src/data_utils.py:183-186) before its float cast."""
    from PIL import Image
    return np.array(Image.open(path).convert('RGB'), dtype=np.uint8)


def load_png16(path):
    """H x W uint16 samples of a 16-bit PNG depth / response map (the raster np.array(Image.open(path)) reads,
    src/data_utils.py:254, 304) before the division by the multiplier."""
    from PIL import Image
    z = np.array(Image.open(path))
    if z.dtype == np.uint8:
        raise ValueError('expected a 16-bit PNG, got 8-bit samples: {}'.format(path))
    return z.astype(np.uint16)


def save_png16(z, path, multiplier):
    """The reference's save_depth / save_response (src/data_utils.py:271-286, 320-335): uint32(z * multiplier) written as a
    Pillow mode-'I' PNG."""
    from PIL import Image
    Image.fromarray(np.uint32(np.asarray(z, dtype=np.float32) * multiplier), mode='I').save(path)


def crop_origin(o_height, o_width, n_height, n_width, crop_type, rng=np.random):
    """(y_start, x_start) of the reference's random_crop (src/datasets.py:19-109), drawing from ``rng`` in the same order
    (horizontal first, then the vertical coin flip and position)."""
    d_height, d_width = o_height - n_height, o_width - n_width
    y_start, x_start = d_height // 2, d_width // 2
    if 'left' in crop_type:
        x_start = 0
    elif 'right' in crop_type:
        x_start = d_width
    elif 'horizontal' in crop_type:
        if 'anchored' in crop_type:
            widths = [a * d_width for a in (0.0, 0.50, 1.0)]
            x_start = int(widths[rng.randint(low=0, high=len(widths))])
        else:
            x_start = rng.randint(low=0, high=d_width)
    if 'top' in crop_type:
        y_start = 0
    elif 'bottom' in crop_type:
        y_start = d_height
    elif 'vertical' in crop_type and rng.rand() <= 0.30:
        if 'anchored' in crop_type:
            heights = [a * d_height for a in (0.0, 0.50, 1.0)]
            y_start = int(heights[rng.randint(low=0, high=len(heights))])
        else:
            y_start = rng.randint(low=0, high=d_height)
    return y_start, x_start


class FusionNetRawDataset(torch.utils.data.Dataset):
    """Same files and crop decisions as the reference's FusionNetTrainingDataset (src/datasets.py:350-440), but the
    sample stays in its on-disk types: (image uint8 H x W x 3, depth / response / ground truth / lidar uint16 H x W,
    crop origin int32 [2]).  ``crop_on_host`` slices the rasters here (fewer bytes over PCIe); otherwise full frames are
    shipped and rcfd_decode_crop crops on the device."""

    def __init__(self, image_paths, depth_paths, response_paths, ground_truth_paths, lidar_map_paths, shape=None,
                 random_crop_type=['none'], crop_on_host=True):
        n = len(image_paths)
        for paths in (depth_paths, response_paths, ground_truth_paths, lidar_map_paths):
            assert len(paths) == n
        self.paths = (image_paths, depth_paths, response_paths, ground_truth_paths, lidar_map_paths)
        self.shape = shape
        self.do_random_crop = shape is not None and all(x > 0 for x in shape)
        self.random_crop_type = random_crop_type
        self.crop_on_host = crop_on_host

    def __len__(self):
        return len(self.paths[0])

    def __getitem__(self, index):
        image = load_image_u8(self.paths[0][index])
        maps = [load_png16(p[index]) for p in self.paths[1:]]
        y0, x0 = 0, 0
        if self.do_random_crop:
            y0, x0 = crop_origin(image.shape[0], image.shape[1], self.shape[0], self.shape[1], self.random_crop_type)
            if self.crop_on_host:
                h, w = self.shape
                image = np.ascontiguousarray(image[y0:y0 + h, x0:x0 + w])
                maps = [np.ascontiguousarray(m[y0:y0 + h, x0:x0 + w]) for m in maps]
                y0, x0 = 0, 0
        return (torch.from_numpy(image),) + tuple(torch.from_numpy(m.view(np.int16)) for m in maps) + \
            (torch.tensor([y0, x0], dtype=torch.int32),)


def decode_fusionnet_batch(raw, device, shape=None, response_multiplier=DEPTH_MULTIPLIER):
    """On-disk sample types -> the five float tensors the reference's DataLoader yields (src/fusionnet_main.py:352-356):
    image N x 3 x H x W in [0, 255], depth, response, ground truth, lidar N x 1 x H x W.  ``raw`` = (image uint8
    N x H0 x W0 x 3, four uint16 N x H0 x W0 maps (int16-viewed tensors are fine), crop origins int32 N x 2); host
    tensors are copied first (non_blocking from pinned memory)."""
    image, depth, response, gt, lidar, origin = [t.to(device, non_blocking=True) for t in raw]
    out_hw = tuple(shape) if shape is not None else tuple(image.shape[1:3])
    crop = origin.contiguous() if out_hw != tuple(image.shape[1:3]) else None
    outs = [ops.decode_crop(image, 1.0, out_hw, crop)]
    for t, mult in ((depth, DEPTH_MULTIPLIER), (response, response_multiplier), (gt, DEPTH_MULTIPLIER),
                    (lidar, DEPTH_MULTIPLIER)):
        outs.append(ops.decode_crop(t, mult, out_hw, crop))
    return outs


def encode_raw_batch(image, depth, response, gt, lidar):
    """float tensors (image N x 3 x H x W in [0, 255]; maps N x 1 x H x W) -> the on-disk sample types of
    FusionNetRawDataset (CPU or CUDA tensors; used to put a synthetic batch into file precision)."""
    n = image.shape[0]
    img = image.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    maps = [(t[:, 0].float() * m).clamp(0, 65535).to(torch.int32).to(torch.uint16).view(torch.int16).contiguous()
            for t, m in ((depth, DEPTH_MULTIPLIER), (response, DEPTH_MULTIPLIER), (gt, DEPTH_MULTIPLIER),
                         (lidar, DEPTH_MULTIPLIER))]
    return [img] + maps + [torch.zeros(n, 2, dtype=torch.int32, device=image.device)]


def make_train_batches(image_path, depth_path, response_path, ground_truth_path, lidar_map_path, batch_size, n_height,
                       n_width, crop_type, n_thread, rank=0, world=1, synthetic_steps=4):
    """Returns (batches(epoch) -> iterator of 5-tuples of CPU tensors, steps per epoch)."""
    if image_path == 'synthetic':
        def batches(epoch):
            for i in range(synthetic_steps):
                seed = 1000 + (epoch * synthetic_steps + i) * world + rank
                image, depth = synth.fusionnet_inputs(batch_size, n_height, n_width, seed)
                gt, lidar = synth.training_targets(batch_size, n_height, n_width, seed)
                yield [t.pin_memory() for t in (image, depth[:, 0:1].contiguous(), depth[:, 1:2].contiguous(), gt, lidar)]
        return batches, synthetic_steps
    paths = [read_paths(p) for p in (image_path, depth_path, response_path, ground_truth_path, lidar_map_path)]
    dataset = FusionNetRawDataset(paths[0], paths[1], paths[2], paths[3], paths[4], shape=(n_height, n_width),
                                  random_crop_type=crop_type)
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world, rank=rank) if world > 1 else None
    loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=sampler is None, sampler=sampler,
                                         num_workers=n_thread, pin_memory=True, drop_last=True)

    def batches(epoch):
        """Raw batches (on-disk sample types): fusionnet_main.train decodes them on the device."""
        if sampler is not None:
            sampler.set_epoch(epoch)
        return iter(loader)
    batches.raw = True
    return batches, len(loader)


def make_val_batches(image_path, depth_path, response_path, ground_truth_path, n_height, n_width, synthetic_samples=4):
    """Validation samples of batch 1 as (image in [0, 255], depth, response, ground_truth) CPU tensors, like the
    reference's FusionNetInferenceDataset behind a DataLoader (reference src/fusionnet_main.py:134-152); None when no
    validation set is given."""
    if image_path is None or image_path == '':
        return None
    if image_path == 'synthetic':
        samples = []
        for i in range(synthetic_samples):
            image, depth = synth.fusionnet_inputs(1, n_height, n_width, 5000 + i)
            gt, _ = synth.training_targets(1, n_height, n_width, 5000 + i)
            samples.append(((image * 255.0).round(), depth[:, 0:1].contiguous(), depth[:, 1:2].contiguous(), gt))
        return samples
    paths = [read_paths(p) for p in (image_path, depth_path, response_path, ground_truth_path)]

    class _Val(torch.utils.data.Dataset):
        """Full frames, no crop (reference FusionNetInferenceDataset, src/datasets.py:443-520), decoded on the host:
        validation is outside the hot path."""

        def __len__(self):
            return len(paths[0])

        def __getitem__(self, i):
            image = torch.from_numpy(load_image_u8(paths[0][i]).astype(np.float32)).permute(2, 0, 1).contiguous()
            maps = []
            for p in paths[1:]:
                z = load_png16(p[i]).astype(np.float32) / DEPTH_MULTIPLIER
                z[z <= 0] = 0.0
                maps.append(torch.from_numpy(z)[None])
            return (image,) + tuple(maps)
    return torch.utils.data.DataLoader(_Val(), batch_size=1, shuffle=False, num_workers=1, drop_last=False)


def make_radarnet_batches(image_path, radar_path, ground_truth_path, batch_size, patch_size, total_points_sampled, n_height,
                          n_width, rank=0, world=1, synthetic_steps=4):
    """Batches of the RadarNet stage-1 training step (reference src/radarnet_main.py:336-340): (image N x 3 x H x (W + pw)
    edge-padded, radar points N x K x 3 with x in padded-image pixels, column boxes N x K x 4, lidar depth of every
    point's crop N x K x 1 x ph x pw).  'synthetic' paths give seeded synthetic batches; the reference's
    nuScenes-derived files are data preparation outside the B200 hot path."""
    if image_path != 'synthetic':
        raise NotImplementedError("RadarNet training data files are outside the B200 hot path: use train_image_path='synthetic'")
    ph, pw = patch_size
    pad = pw // 2
    k = total_points_sampled

    def batches(epoch):
        for i in range(synthetic_steps):
            seed = 3000 + i * world + rank              # the same few synthetic batches every epoch
            g = torch.Generator().manual_seed(seed)
            image = torch.rand(batch_size, 3, n_height, n_width + 2 * pad, generator=g)
            pts = torch.stack([synth.radar_points(k, n_height, n_width, seed * 131 + b) for b in range(batch_size)])
            pts[..., 0] += pad
            boxes = torch.stack([pts[..., 0] - pad, torch.zeros(batch_size, k), pts[..., 0] + pad,
                                 torch.full((batch_size, k), float(n_height))], dim=-1)
            # sparse lidar returns in every crop; a band of them lies within the correspondence distance of the radar depth
            z = pts[..., 2].view(batch_size, k, 1, 1, 1)
            noise = (torch.rand(batch_size, k, 1, ph, pw, generator=g) - 0.5) * 6.0
            far = torch.rand(batch_size, k, 1, ph, pw, generator=g) * 79 + 1
            near = torch.rand(batch_size, k, 1, ph, pw, generator=g) < 0.3
            mask = torch.rand(batch_size, k, 1, ph, pw, generator=g) < 0.05
            gt = torch.where(near, (z + noise).clamp_min(0.5), far) * mask
            yield [t.pin_memory() for t in (image, pts, boxes, gt.float())]
    return batches, synthetic_steps
