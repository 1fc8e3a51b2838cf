"""
MEASUREMENT COMPARATOR (not product code, not the oracle): the reference FusionNet graph written with the same
ATen calls its torch.nn modules make (F.conv2d bias-free + F.batch_norm + F.leaky_relu, F.max_pool2d,
F.interpolate(nearest), torch.cat: reference src/net_utils.py:63-91, 156-198, 253-323, 473-569,
src/networks.py:840-1005, 1557-1657), so that `bench.py --impl cudnn` can time what the reference itself would run
on this GPU -- PyTorch eager + cuDNN, which ships sm_100 kernels -- in fp32 (TF32 allowed: torch's cuDNN default)
and under autocast(bf16) + channels_last.  Parameters come from the product model's state_dict (same names), the
training step is forward + outlier removal + masked L1 + backward + torch.optim.Adam like src/fusionnet_main.py:348-399.
None of the repo's kernels are on this path.
"""
import torch
import torch.nn.functional as F

SLOPE = 0.2


def conv(p, pre, x, stride, act, training, bn=True):
    w = p[pre + '.conv.weight']
    y = F.conv2d(x, w, None, stride, w.shape[-1] // 2)
    if bn:
        y = F.batch_norm(y, p[pre + '.batch_norm.running_mean'], p[pre + '.batch_norm.running_var'],
                         p[pre + '.batch_norm.weight'], p[pre + '.batch_norm.bias'], training, 0.1, 1e-5)
    if act == 'leaky':
        y = F.leaky_relu(y, SLOPE)
    elif act == 'sigmoid':
        y = torch.sigmoid(y)
    return y


def block(p, pre, x, stride, training):
    c1 = conv(p, pre + '.conv1', x, stride, 'leaky', training)
    c2 = conv(p, pre + '.conv2', c1, 1, 'leaky', training)
    sc = x if x.shape[1:] == c2.shape[1:] else conv(p, pre + '.projection', x, stride, None, training, bn=False)
    return F.leaky_relu(c2 + sc, SLOPE)


def stage(p, pre, x, stride, training):
    x = block(p, pre + '.0', x, stride, training)
    return block(p, pre + '.1', x, 1, training)


def forward(p, image, depth, training, min_depth=1.0, max_depth=100.0):
    e = 'encoder.'
    ci = conv(p, e + 'conv1_image', image, 2, 'leaky', training)
    cd = conv(p, e + 'conv1_depth', depth, 2, 'leaky', training)

    def fuse(level, img, dep):
        return conv(p, e + 'conv%d_weight' % level, dep, 1, 'sigmoid', training) * \
            conv(p, e + 'conv%d_project' % level, dep, 1, None, training) + img
    layers = [fuse(1, ci, cd)]
    xi, xd = F.max_pool2d(ci, 3, 2, 1), F.max_pool2d(cd, 3, 2, 1)
    for level in range(2, 7):
        s = 1 if level == 2 else 2
        xi = stage(p, e + 'blocks%d_image' % level, xi, s, training)
        xd = stage(p, e + 'blocks%d_depth' % level, xd, s, training)
        layers.append(fuse(level, xi, xd))
    x, skips = layers[-1], layers[:-1]
    for b in range(5, -1, -1):
        skip = skips[b - 1] if b >= 1 else None
        size = skip.shape[2:4] if skip is not None else image.shape[2:4]
        d = conv(p, 'decoder.deconv%d.deconv.conv' % b, F.interpolate(x, size=tuple(size)), 1, 'leaky', training)
        if skip is not None:
            d = torch.cat([d, skip], 1)
        x = conv(p, 'decoder.deconv%d.conv' % b, d, 1, 'leaky', training)
    logits = conv(p, 'decoder.output0', x, 1, None, training, bn=False)
    return min_depth / (torch.sigmoid(logits.float()) + min_depth / max_depth)


def outlier_removal(depth, k=7, threshold=1.5):
    mx = 10 * torch.max(depth)
    filled = torch.where(depth > 0, depth, mx.expand_as(depth))
    pad = k // 2
    filled = F.pad(filled, (pad, pad, pad, pad), value=float(mx))          # .item() sync like the reference (:619)
    mn = -F.max_pool2d(-filled, k, 1, 0)
    return torch.where(mn < depth - threshold, torch.zeros_like(depth), depth)


def loss_fn(out, gt, lidar, w_lidar=2.0):
    gt = gt * (lidar <= 0).float()
    vg, vl = gt > 0, lidar > 0
    return F.l1_loss(out[vg], gt[vg]) + w_lidar * F.l1_loss(out[vl], lidar[vl])          # boolean-mask gathers (:245-253)


def make_step(state, mode, batch_tensors, variant):
    """state: name -> CUDA tensor (product state_dict with encoder./decoder. prefixes).  variant: 'fp32' (TF32 conv
    allowed, torch default) or 'bf16_cl' (autocast bf16 + channels_last).  Returns a zero-argument step function."""
    image, depth, gt, lidar = batch_tensors
    p = {k: v.detach().clone() for k, v in state.items()}
    cl = variant == 'bf16_cl'
    if cl:
        image = image.contiguous(memory_format=torch.channels_last)
        depth = depth.contiguous(memory_format=torch.channels_last)
        for k in p:
            if p[k].dim() == 4:
                p[k] = p[k].contiguous(memory_format=torch.channels_last)
    ctx = (lambda: torch.autocast('cuda', dtype=torch.bfloat16)) if cl else (lambda: torch.autocast('cuda', enabled=False))
    if mode == 'infer':
        def step():
            with torch.no_grad(), ctx():
                return forward(p, image, depth, False)
        return step
    names = [k for k, v in p.items() if v.is_floating_point() and 'running' not in k]
    for k in names:
        p[k].requires_grad_(True)
    opt = torch.optim.Adam([p[k] for k in names], lr=1e-3)

    def step():
        with ctx():
            out = forward(p, image, depth, True)
        loss = loss_fn(out.float(), outlier_removal(gt), lidar)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.detach()
    return step
