"""
TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's on-device augmentation
`fusionnet_transforms.Transforms.transform` (reference src/fusionnet_transforms.py:46-178) -- SURVEY.md 8f row 2,
prepared for the batched CUDA kernel of the next round.  Never imported by the product.

The reference draws its per-sample decisions with consecutive `torch.rand(n_batch)` calls and then loops over the
samples calling torchvision's functional ops.  Two parts here:

  * draw_decisions(...)   the SAME sequence of torch.rand calls (so a seeded generator state reproduces the
                          reference's choices): which samples get brightness / contrast / saturation / flips, and
                          the factors (reference :76-121, :139-165);
  * apply(...)            the arithmetic, batched, with the exact semantics the reference inherits from
                          torchvision 0.26 `_functional_tensor` for the dtype it feeds it:
      - images whose maximum exceeds 1.0 are cast with `.int()` (int32!) first (reference :80-83), so every blend is
        `trunc(clamp(f * img + (1 - f) * other, 0, 2**31 - 1))` -- NO clamp at 255, truncation toward zero after each
        op; float images in [0, 1] clamp to [0, 1] and do not truncate;
      - brightness  : other = 0                                   (torchvision adjust_brightness -> _blend)
      - contrast    : other = mean over the whole image of gray, gray = (0.2989 r + 0.587 g + 0.114 b) cast to the
                      image dtype (truncated for int32), mean taken in float32
      - saturation  : other = gray (per pixel, image dtype)
      - then `.float()`, normalisation to [0, 1] / [-1, 1] / [0, 255] (reference :126-136, :186-213), then the flips
        of images AND range maps (reference :141-165).
    fp32 products and sums are rounded separately (no fused multiply-add), in the order `f * img + (1 - f) * other`.
"""
import torch


def draw_decisions(n_batch, probability, random_brightness=(-1,), random_contrast=(-1,), random_saturation=(-1,),
                   random_flip_type=('none',), generator=None, device='cpu'):
    """Replays the reference's random draws (same number and order of torch.rand(n_batch) calls)."""
    rand = lambda: torch.rand(n_batch, device=device, generator=generator)
    d = {'n': n_batch}
    do_t = rand() <= probability                                        # reference :76-77
    for name, cfg in (('brightness', random_brightness), ('contrast', random_contrast),
                      ('saturation', random_saturation)):
        if -1 in cfg:
            continue
        do = torch.logical_and(do_t, rand() <= 0.50)                    # :87-89 / :100-102 / :113-115
        values = rand()
        lo, hi = cfg
        d[name] = (do, (hi - lo) * values + lo)
    if 'horizontal' in random_flip_type:
        d['hflip'] = torch.logical_and(do_t, rand() <= 0.50)            # :141-143
    if 'vertical' in random_flip_type:
        d['vflip'] = torch.logical_and(do_t, rand() <= 0.50)            # :154-156
    return d


def _gray(img):
    r, g, b = img.unbind(dim=-3)
    return (0.2989 * r + 0.587 * g + 0.114 * b).to(img.dtype).unsqueeze(-3)


def _blend(img, other, f):
    """torchvision _blend for a batch: f is N floats (python-float precision in the reference: float(ratio))."""
    bound = 1.0 if img.is_floating_point() else 2147483647
    f = f.view(-1, 1, 1, 1).double()                                     # the reference multiplies by a python float
    a = (f.float() * img)                                                # float32 product (weak-scalar promotion)
    b = ((1.0 - f).float() * other)
    return (a + b).clamp(0, bound).to(img.dtype)


def apply(images, range_maps, decisions, normalized_image_range=(0, 1)):
    """images: N x 3 x H x W float (0..255 or 0..1); range_maps: list of N x c x H x W.  Returns (images, range_maps)."""
    images = images.clone()
    if torch.max(images) > 1.0:
        images = images.int()                                            # reference :80-83
    n = images.shape[0]
    for name in ('brightness', 'contrast', 'saturation'):
        if name not in decisions:
            continue
        do, factors = decisions[name]
        if name == 'brightness':
            other = torch.zeros_like(images)
        elif name == 'contrast':
            other = torch.mean(_gray(images).to(torch.float32 if not images.is_floating_point() else images.dtype),
                               dim=(-3, -2, -1), keepdim=True)
        else:
            other = _gray(images)
        blended = _blend(images, other, factors.to(images.device))
        sel = do.view(n, 1, 1, 1).to(images.device)
        images = torch.where(sel, blended, images)
    images = images.float()
    rng = list(normalized_image_range)
    if rng == [0, 1]:
        images = images / 255.0
    elif rng == [-1, 1]:
        images = 2.0 * (images / 255.0) - 1.0
    elif rng != [0, 255]:
        raise ValueError('Unsupported normalization range: {}'.format(rng))
    range_maps = [m.clone() for m in range_maps]
    for key, dim in (('hflip', -1), ('vflip', -2)):
        if key in decisions:
            sel = decisions[key].view(n, 1, 1, 1)
            images = torch.where(sel.to(images.device), torch.flip(images, dims=[dim]), images)
            range_maps = [torch.where(sel.to(m.device), torch.flip(m, dims=[dim]), m) for m in range_maps]
    return images, range_maps
