"""
Transforms with the reference's constructor / `transform` surface (reference: src/fusionnet_transforms.py:4-334),
batched: the reference loops over the samples in Python and calls torchvision's functional ops on each; here every
step is one tensor expression over the whole batch on whatever device the tensors live on.  Same random draws (the
same sequence of `torch.rand(n_batch)` calls), same arithmetic, bit for bit (tests/golden/transforms_5x18x26.npz was
written by the reference itself):

  * images whose maximum exceeds 1.0 are cast with `.int()` first (reference :80-83), so -- exactly as torchvision
    0.26 treats int32 tensors -- every blend is trunc(clamp(f * img + (1 - f) * other, 0, 2**31 - 1)): no clamp at 255;
    float images in [0, 1] clamp to [0, 1];
  * brightness blends with 0, contrast with the mean of the grey image (grey = 0.2989 r + 0.587 g + 0.114 b cast to the
    image dtype), saturation with the grey image;
  * normalisation to [0, 1] / [-1, 1] / [0, 255], then horizontal / vertical flips of images AND range maps.

CUDA tensors go through ONE fused apply kernel for the whole batch (rcfd_transform_batch: blends + normalisation +
flips of the image and of every range map, preceded by a max and a per-sample grey-mean reduction; no host
synchronisation, unlike the reference's `torch.max(images) > 1.0` at :82).  CPU tensors take the equivalent tensor
expressions below (host-side tests).  The only arithmetic difference between the two: the contrast partner (mean of the
grey image) is summed exactly in float64 by the kernel and in float32 by torch.mean, which can move a blended value
across an integer boundary on isolated pixels (bounded by one grey level; tests/test_dataops_gpu.py).
"""
import torch

from rcfd import ops


class Transforms(object):

    def __init__(self, normalized_image_range=[0, 255], random_brightness=[-1], random_contrast=[-1],
                 random_saturation=[-1], random_flip_type=['none'], rand_device=None):
        self.normalized_image_range = normalized_image_range
        self.rand_device = rand_device          # where the torch.rand draws are made (None: the data's device)
        self.do_random_brightness = -1 not in random_brightness
        self.random_brightness = random_brightness
        self.do_random_contrast = -1 not in random_contrast
        self.random_contrast = random_contrast
        self.do_random_saturation = -1 not in random_saturation
        self.random_saturation = random_saturation
        self.do_random_horizontal_flip = 'horizontal' in random_flip_type
        self.do_random_vertical_flip = 'vertical' in random_flip_type

    # ------------------------------------------------------------------ batched building blocks
    @staticmethod
    def _gray(img):
        r, g, b = img.unbind(dim=-3)
        return (0.2989 * r + 0.587 * g + 0.114 * b).to(img.dtype).unsqueeze(-3)

    @staticmethod
    def _blend(img, other, factors, do):
        """torchvision `_blend` for the selected samples: ratio * img + (1 - ratio) * other, clamped, cast back."""
        bound = 1.0 if img.is_floating_point() else 2147483647
        f = factors.view(-1, 1, 1, 1).double()           # the reference passes python floats: (1 - f) in double
        blended = (f.float() * img + (1.0 - f).float() * other).clamp(0, bound).to(img.dtype)
        return torch.where(do.view(-1, 1, 1, 1), blended, img)

    def transform(self, images_arr, range_maps_arr=[], random_transform_probability=0.50):
        """list of N x C x H x W images (+ list of N x c x H x W range maps) -> the same lists, transformed
        (reference :46-178: same return convention)."""
        device = images_arr[0].device
        if images_arr[0].ndim != 4:
            raise ValueError('Unsupported number of dimensions: {}'.format(images_arr[0].ndim))
        n_batch = images_arr[0].shape[0]
        images_arr = list(images_arr)
        range_maps_arr = list(range_maps_arr)
        rand_device = device if self.rand_device is None else self.rand_device
        rand = lambda: torch.rand(n_batch, device=rand_device)
        if device.type == 'cuda':
            return self._transform_cuda(images_arr, range_maps_arr, random_transform_probability, rand, device)
        do_random_transform = rand() <= random_transform_probability

        for idx, images in enumerate(images_arr):
            if torch.max(images) > 1.0:                  # [0, 255] images passed as float (reference :80-83)
                images_arr[idx] = images.int()

        for enabled, limits, kind in ((self.do_random_brightness, self.random_brightness, 'brightness'),
                                      (self.do_random_contrast, self.random_contrast, 'contrast'),
                                      (self.do_random_saturation, self.random_saturation, 'saturation')):
            if not enabled:
                continue
            do = torch.logical_and(do_random_transform, rand() <= 0.50)
            values = rand()
            lo, hi = limits
            factors = (hi - lo) * values + lo
            for idx, images in enumerate(images_arr):
                if kind == 'brightness':
                    other = torch.zeros_like(images)
                elif kind == 'contrast':
                    dtype = images.dtype if images.is_floating_point() else torch.float32
                    other = torch.mean(self._gray(images).to(dtype), dim=(-3, -2, -1), keepdim=True)
                else:
                    other = self._gray(images)
                images_arr[idx] = self._blend(images, other, factors, do)

        images_arr = [images.float() for images in images_arr]
        images_arr = self.normalize_images(images_arr, normalized_image_range=self.normalized_image_range)

        for enabled, dim in ((self.do_random_horizontal_flip, -1), (self.do_random_vertical_flip, -2)):
            if not enabled:
                continue
            do = torch.logical_and(do_random_transform, rand() <= 0.50).view(-1, 1, 1, 1)
            images_arr = [torch.where(do, torch.flip(t, dims=[dim]), t) for t in images_arr]
            range_maps_arr = [torch.where(do, torch.flip(t, dims=[dim]), t) for t in range_maps_arr]

        outputs = []
        if len(images_arr) > 0:
            outputs.append(images_arr)
        if len(range_maps_arr) > 0:
            outputs.append(range_maps_arr)
        return outputs[0] if len(outputs) == 1 else outputs

    def _transform_cuda(self, images_arr, range_maps_arr, probability, rand, device):
        """Same draws (same sequence of torch.rand(n_batch) calls as the reference), arithmetic in rcfd_transform_batch."""
        n_batch = images_arr[0].shape[0]
        do_t = rand() <= probability
        zeros = torch.zeros(n_batch, device=do_t.device)
        cols = []
        for enabled, limits in ((self.do_random_brightness, self.random_brightness),
                                (self.do_random_contrast, self.random_contrast),
                                (self.do_random_saturation, self.random_saturation)):
            if not enabled:
                cols += [zeros, zeros, zeros]
                continue
            do = torch.logical_and(do_t, rand() <= 0.50)
            lo, hi = limits
            factors = (hi - lo) * rand() + lo
            cols += [do.float(), factors, (1.0 - factors.double()).float()]      # (1 - f) in double like the python float
        for enabled in (self.do_random_horizontal_flip, self.do_random_vertical_flip):
            cols.append(torch.logical_and(do_t, rand() <= 0.50).float() if enabled else zeros)
        params = torch.stack(cols, dim=1).to(device=device, dtype=torch.float32)
        rng = list(self.normalized_image_range)
        if rng not in ([0, 255], [0, 1], [-1, 1]):
            raise ValueError('Unsupported normalization range: {}'.format(self.normalized_image_range))
        norm_mode = {(0, 255): 0, (0, 1): 1, (-1, 1): 2}[tuple(rng)]
        images_out, maps_out = [], [m.float() for m in range_maps_arr]
        for idx, images in enumerate(images_arr):
            img, maps = ops.transform_batch(images.float(), maps_out if idx == 0 else [], params, norm_mode)
            images_out.append(img)
            if idx == 0:
                maps_out = maps
        if len(images_arr) == 0 and maps_out:
            _, maps_out = ops.transform_batch(None, maps_out, params, norm_mode)
        outputs = []
        if len(images_out) > 0:
            outputs.append(images_out)
        if len(maps_out) > 0:
            outputs.append(maps_out)
        return outputs[0] if len(outputs) == 1 else outputs

    def normalize_images(self, images_arr, normalized_image_range=[0, 1]):
        if normalized_image_range == [0, 1]:
            return [images / 255.0 for images in images_arr]
        if normalized_image_range == [-1, 1]:
            return [2.0 * (images / 255.0) - 1.0 for images in images_arr]
        if normalized_image_range == [0, 255]:
            return images_arr
        raise ValueError('Unsupported normalization range: {}'.format(normalized_image_range))
