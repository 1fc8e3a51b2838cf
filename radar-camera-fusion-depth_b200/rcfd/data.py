"""
Batch sources for fusionnet_main.train.  Dataset file formats are out of scope for the accelerated
path: real data is read by the REFERENCE's own
``datasets.FusionNetTrainingDataset`` / ``data_utils.read_paths`` when its ``src`` directory is on
``sys.path``; ``'synthetic'`` paths produce the seeded synthetic batches of SURVEY 8d.
"""
import torch

from . import synth


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
    try:
        import datasets                # the reference's src/datasets.py
        import data_utils              # the reference's src/data_utils.py
    except ImportError as e:
        raise ImportError("reading nuScenes-derived training data needs the reference's datasets.py / data_utils.py on "
                          "sys.path (file formats are outside the B200 hot path); use train_image_path='synthetic' for "
                          "the synthetic workload") from e
    paths = [data_utils.read_paths(p) for p in (image_path, depth_path, response_path, ground_truth_path, lidar_map_path)]
    dataset = datasets.FusionNetTrainingDataset(
        image_paths=paths[0], depth_paths=paths[1], response_paths=paths[2], ground_truth_paths=paths[3],
        lidar_map_paths=paths[4], shape=(n_height, n_width), random_crop_type=crop_type)
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world, rank=rank) if world > 1 else None
    loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=sampler is None, sampler=sampler,
                                         num_workers=n_thread, pin_memory=True, drop_last=True)

    def batches(epoch):
        if sampler is not None:
            sampler.set_epoch(epoch)
        return iter(loader)
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
    try:
        import datasets
        import data_utils
    except ImportError as e:
        raise ImportError("reading nuScenes-derived validation data needs the reference's datasets.py / data_utils.py on "
                          "sys.path; use val_image_path='synthetic' or '' otherwise") from e
    paths = [data_utils.read_paths(p) for p in (image_path, depth_path, response_path, ground_truth_path)]
    dataset = datasets.FusionNetInferenceDataset(image_paths=paths[0], depth_paths=paths[1], response_paths=paths[2],
                                                 ground_truth_paths=paths[3])
    return torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=False, num_workers=1, drop_last=False)


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
