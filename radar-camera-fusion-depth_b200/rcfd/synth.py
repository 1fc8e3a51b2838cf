"""
Deterministic synthetic weights and inputs.

Everything is drawn from seeded CPU ``torch.Generator``s so the CPU oracle, the golden
fixtures and the GPU path all see identical bits.  No dataset, no checkpoint.
"""
import math
import torch

CANONICAL_FUSIONNET = dict(
    input_channels_image=3,
    input_channels_depth=2,
    encoder_type=['fusionnet18', 'batch_norm'],
    n_filters_encoder_image=[32, 64, 128, 256, 256, 256],
    n_filters_encoder_depth=[16, 32, 64, 128, 128, 128],
    fusion_type='weight_and_project',
    decoder_type=['multiscale', 'batch_norm'],
    n_resolution_decoder=1,
    n_filters_decoder=[256, 256, 128, 64, 64, 32],
    deconv_type='up',
    activation_func='leaky_relu',
    weight_initializer='kaiming_uniform',
    min_predict_depth=1.0,
    max_predict_depth=100.0)

# A narrow variant of the same topology, small enough to ship golden weights for.
SMALL_FUSIONNET = dict(CANONICAL_FUSIONNET,
                       n_filters_encoder_image=[16, 16, 32, 32, 32, 32],
                       n_filters_encoder_depth=[8, 16, 16, 16, 16, 16],
                       n_filters_decoder=[32, 32, 32, 16, 16, 16])

CANONICAL_RADARNET = dict(
    input_channels_image=3,
    input_channels_depth=3,
    input_patch_size_image=(352, 288),
    encoder_type=['radarnetv1', 'batch_norm'],
    n_filters_encoder_image=[32, 64, 128, 128, 128],
    n_neurons_encoder_depth=[32, 64, 128, 128, 128],
    decoder_type=['multiscale', 'batch_norm'],
    n_filters_decoder=[256, 128, 64, 32, 16],
    weight_initializer='kaiming_uniform',
    activation_func='leaky_relu')


def _gen(seed):
    g = torch.Generator(device='cpu')
    g.manual_seed(int(seed))
    return g


def fill_state_dict_(state, seed=0):
    """Overwrite every tensor of a ``state_dict``-like mapping (iterated in its own
    order) with deterministic values: fan-in scaled uniform conv / linear weights and
    non-trivial BatchNorm affine + running statistics."""
    for idx, (key, t) in enumerate(state.items()):
        g = _gen(seed * 1000003 + idx)
        if key.endswith('num_batches_tracked'):
            t.zero_()
        elif 'batch_norm' in key and key.endswith('weight'):
            t.copy_(torch.rand(t.shape, generator=g) + 0.5)
        elif 'batch_norm' in key and key.endswith('bias'):
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif key.endswith('running_mean'):
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif key.endswith('running_var'):
            t.copy_(torch.rand(t.shape, generator=g) + 0.5)
        elif key.endswith('weight') and t.dim() >= 2:
            fan_in = t[0].numel()
            bound = math.sqrt(3.0 / fan_in)          # unit-gain: keeps eval-mode activations O(1)
            if 'conv1_depth.conv' in key:
                bound *= 0.03                        # input depth is in metres (up to 80)
            t.copy_((torch.rand(t.shape, generator=g) * 2 - 1) * bound)
        else:   # linear bias
            t.copy_((torch.rand(t.shape, generator=g) * 2 - 1) * 0.1)
    return state


def radar_points(k, h, w, seed=0):
    """K radar returns (x, y, z); the first 8 exercise rounding ties, duplicates and
    the legal border."""
    g = _gen(seed + 77)
    x = torch.rand(k, generator=g) * (w - 2) + 1
    y = torch.rand(k, generator=g) * (h - 2) + 1
    z = torch.rand(k, generator=g) * 79 + 1
    if k >= 8:
        x[0], y[0] = 10.5, 20.5          # ties: half-to-even -> (10, 20)
        x[1], y[1] = 11.5, 21.5          # -> (12, 22)
        x[2], y[2] = 30.5, 7.5
        x[3], y[3] = 31.5, 8.5
        x[4], y[4] = x[5], y[5]          # exact duplicates (last writer wins)
        x[6], y[6] = 1.25, 1.25          # legal border (1 < x)
        x[7], y[7] = w - 1.5, h - 1.5
    return torch.stack([x, y, z], dim=1)


def _scatter_sparse(points, h, w):
    img = torch.zeros(h, w)
    ix = torch.round(points[:, 0]).long().clamp_(0, w - 1)   # torch.round is half-to-even
    iy = torch.round(points[:, 1]).long().clamp_(0, h - 1)
    for i in range(points.shape[0]):
        img[iy[i], ix[i]] = points[i, 2]
    return img


def _scatter_bands(points, h, w):
    img = torch.zeros(h, w)
    for i in range(points.shape[0]):
        x, y, z = [float(v) for v in points[i]]
        x0, x1 = max(0, int(math.ceil(x - 16))), min(w, int(math.floor(x + 16)) + 1)
        y0, y1 = max(0, int(y) - 48), min(h, int(y) + 16)
        if x0 >= x1 or y0 >= y1:
            continue
        band = img[y0:y1, x0:x1]
        img[y0:y1, x0:x1] = torch.where((band == 0) | (band > z), torch.full_like(band, z), band)
    return img


def fusionnet_inputs(n, h, w, seed=0, variant='quasi_dense', k=64):
    """image U[0,1) N x 3 x H x W; input_depth N x 2 x H x W = (depth, response)."""
    g = _gen(seed)
    image = torch.rand(n, 3, h, w, generator=g)
    depth = torch.zeros(n, 2, h, w)
    for b in range(n):
        pts = radar_points(k, h, w, seed * 131 + b)
        d = _scatter_sparse(pts, h, w) if variant == 'sparse' else _scatter_bands(pts, h, w)
        r = torch.rand(h, w, generator=g) * 0.5 + 0.5
        depth[b, 0] = d
        depth[b, 1] = torch.where(d > 0, r, torch.zeros_like(r))
    return image, depth


def training_targets(n, h, w, seed=0):
    """ground_truth: U[1,80) on a Bernoulli(0.30) mask; lidar_map: U[1,80) on Bernoulli(0.02)."""
    g = _gen(seed + 991)
    gt = (torch.rand(n, 1, h, w, generator=g) * 79 + 1) * (torch.rand(n, 1, h, w, generator=g) < 0.30)
    li = (torch.rand(n, 1, h, w, generator=g) * 79 + 1) * (torch.rand(n, 1, h, w, generator=g) < 0.02)
    return gt, li
