"""
ORACLE (test infrastructure, NOT product code) -- CPU fp32 restatement of the
reference FusionNet hot path in plain functional PyTorch.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path (radar-camera-fusion-depth_b200/) never
does: it fails loudly if the CUDA extension is missing.

Parity status: PINNED against the unmodified reference executed in the build
container (tests/golden/make_golden.py imports /root/reference/src and writes
tests/golden/*.npz; tests/test_oracle_golden.py replays them).  The reference
itself ships no tests / golden vectors, so those fixtures are
the pin.

Every function cites the reference file:line it restates (paths relative to the
reference repository root).  Tensors are NCHW fp32 like the reference; parameters
come in as a flat dict keyed exactly like the reference ``state_dict`` with an
``encoder.`` / ``decoder.`` prefix.
"""
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.20      # src/net_utils.py:15  (activation_func('leaky_relu'))
BN_EPS = 1e-5           # torch.nn.BatchNorm2d default, src/net_utils.py:82
BN_MOMENTUM = 0.1


def _act(x, kind):
    # src/net_utils.py:4-23
    if kind is None or kind == 'linear':
        return x
    if kind == 'leaky_relu':
        return F.leaky_relu(x, LEAKY_SLOPE)
    if kind == 'sigmoid':
        return torch.sigmoid(x)
    raise ValueError(kind)


def conv_block(p, prefix, x, stride, act, use_bn, training, new_stats=None):
    """src/net_utils.py:29-91  Conv2d: conv(bias=False, pad=k//2) -> BN? -> act?"""
    w = p[prefix + '.conv.weight']
    k = w.shape[-1]
    y = F.conv2d(x, w, None, stride=stride, padding=k // 2)
    if use_bn:
        rm = p[prefix + '.batch_norm.running_mean']
        rv = p[prefix + '.batch_norm.running_var']
        g = p[prefix + '.batch_norm.weight']
        b = p[prefix + '.batch_norm.bias']
        if training:
            # batch statistics, biased variance for normalisation,
            # unbiased for the running estimate (torch.nn.BatchNorm2d semantics)
            mean = y.mean(dim=(0, 2, 3))
            var = y.var(dim=(0, 2, 3), unbiased=False)
            if new_stats is not None:
                n = y.numel() // y.shape[1]
                with torch.no_grad():
                    new_stats[prefix + '.batch_norm.running_mean'] = \
                        (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                    new_stats[prefix + '.batch_norm.running_var'] = \
                        (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * (n / max(n - 1, 1))
        else:
            mean, var = rm, rv
        y = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
        y = y * g[None, :, None, None] + b[None, :, None, None]
    return _act(y, act)


def resnet_block(p, prefix, x, stride, use_bn, training, new_stats=None):
    """src/net_utils.py:253-323.  Note act is applied to conv2 BEFORE the residual
    add and again after it (:291-298, :323); projection has no BN / act (:300-307)
    and is used only when shape or channel count changes (:315-320)."""
    c1 = conv_block(p, prefix + '.conv1', x, stride, 'leaky_relu', use_bn, training, new_stats)
    c2 = conv_block(p, prefix + '.conv2', c1, 1, 'leaky_relu', use_bn, training, new_stats)
    if list(x.shape[1:]) != list(c2.shape[1:]):
        sc = conv_block(p, prefix + '.projection', x, stride, None, False, training)
    else:
        sc = x
    return F.leaky_relu(c2 + sc, LEAKY_SLOPE)


def resnet_stage(p, prefix, x, stride, n_block, use_bn, training, new_stats=None):
    """src/networks.py:178-230 / :767-838 (_make_layer): first block strided."""
    for n in range(n_block):
        x = resnet_block(p, '%s.%d' % (prefix, n), x, stride if n == 0 else 1,
                         use_bn, training, new_stats)
    return x


def fusionnet_encoder(p, image, depth, n_levels=6, use_bn=True, training=False,
                      new_stats=None, taps=None, pre='encoder.'):
    """src/networks.py:840-1005, fusion_type='weight_and_project' (:863-866 ...).
    The UNFUSED branch tensors feed the next level (:875-879, :899-900)."""
    ci = conv_block(p, pre + 'conv1_image', image, 2, 'leaky_relu', use_bn, training, new_stats)
    cd = conv_block(p, pre + 'conv1_depth', depth, 2, 'leaky_relu', use_bn, training, new_stats)

    def fuse(level, img, dep):
        w = conv_block(p, pre + 'conv%d_weight' % level, dep, 1, 'sigmoid', use_bn, training, new_stats)
        pr = conv_block(p, pre + 'conv%d_project' % level, dep, 1, None, use_bn, training, new_stats)
        return w * pr + img

    layers = [fuse(1, ci, cd)]
    xi = F.max_pool2d(ci, 3, 2, 1)      # src/networks.py:392-395, :875-876
    xd = F.max_pool2d(cd, 3, 2, 1)
    for level in range(2, n_levels + 1):
        s = 1 if level == 2 else 2
        xi = resnet_stage(p, pre + 'blocks%d_image' % level, xi, s, 2, use_bn, training, new_stats)
        xd = resnet_stage(p, pre + 'blocks%d_depth' % level, xd, s, 2, use_bn, training, new_stats)
        layers.append(fuse(level, xi, xd))
    if taps is not None:
        taps['conv1_image'] = ci
        taps['conv1_depth'] = cd
        taps['blocks_image_last'] = xi
        taps['blocks_depth_last'] = xd
    return layers[-1], layers[:-1]


def decoder_block(p, prefix, x, skip, shape, use_bn, training, new_stats=None):
    """src/net_utils.py:473-569 with deconv_type='up' (:156-198): nearest
    interpolate to skip's (or the given) size -> 3x3 conv -> cat skip -> 3x3 conv."""
    if skip is not None:
        shape = skip.shape[2:4]
    up = F.interpolate(x, size=tuple(shape))
    d = conv_block(p, prefix + '.deconv.conv', up, 1, 'leaky_relu', use_bn, training, new_stats)
    if skip is not None:
        d = torch.cat([d, skip], dim=1)
    return conv_block(p, prefix + '.conv', d, 1, 'leaky_relu', use_bn, training, new_stats)


def multiscale_decoder(p, latent, skips, shape, use_bn=True, training=False,
                       new_stats=None, taps=None, pre='decoder.', all_outputs=False):
    """src/networks.py:1557-1657.  Decoder depth follows the number of filters: deconv{len-1} ... deconv0, the last one
    without a skip (:1647-1652) when there are fewer skips than blocks.  n_resolution > 1 (detected from the output1..3
    weights): after deconv{b} (b = 3, 2, 1 when n_resolution > b) a 3x3 output conv gives the logits at that resolution
    (:1595-1598, :1610-1613, :1626-1629), which are up-sampled 2x bilinearly with align_corners=True (:1600-1604, ...) and
    concatenated BEHIND the next block's skip (:1608, :1624, :1640-1642; alone when that block has no skip).
    Returns the full-resolution logits, or with all_outputs the list [coarsest ..., output0] (:1656)."""
    n_blocks = 0
    while (pre + 'deconv%d.conv.conv.weight' % n_blocks) in p:
        n_blocks += 1
    n_resolution = 1 + sum(1 for b in (1, 2, 3) if (pre + 'output%d.conv.weight' % b) in p)
    x = latent
    n = len(skips) - 1
    outputs, up = [], None
    for b in range(n_blocks - 1, -1, -1):
        skip = skips[n] if n >= 0 else None
        n -= 1
        if up is not None:
            skip = torch.cat([skip, up], dim=1) if skip is not None else up
        x = decoder_block(p, pre + 'deconv%d' % b, x, skip, None if skip is not None else shape, use_bn, training, new_stats)
        if taps is not None:
            taps['deconv%d' % b] = x
        up = None
        if 1 <= b <= 3 and n_resolution > b:
            out_b = conv_block(p, pre + 'output%d' % b, x, 1, None, False, training)
            outputs.append(out_b)
            up = F.interpolate(out_b, scale_factor=2, mode='bilinear', align_corners=True)
    outputs.append(conv_block(p, pre + 'output0', x, 1, None, False, training))
    return outputs if all_outputs else outputs[-1]


def fusionnet_forward(p, image, input_depth, min_predict_depth=1.0, max_predict_depth=100.0,
                      n_levels=6, training=False, new_stats=None, taps=None, return_multiscale=False):
    """src/fusionnet_model.py:140-170.  Returns (depth, logits); with return_multiscale both are lists over the decoder's
    output resolutions, coarsest first."""
    latent, skips = fusionnet_encoder(p, image, input_depth, n_levels, True, training, new_stats, taps)
    if taps is not None:
        taps['latent'] = latent
        for i, s in enumerate(skips):
            taps['skip%d' % (i + 1)] = s
    logits = multiscale_decoder(p, latent, skips, image.shape[-2:], True, training, new_stats, taps, all_outputs=True)
    depth = [min_predict_depth / (torch.sigmoid(o) + min_predict_depth / max_predict_depth) for o in logits]
    if return_multiscale:
        return depth, logits
    return depth[-1], logits[-1]


def fusionnet_loss(output_depth, ground_truth, lidar_map, w_lidar_loss=2.0, loss_func='l1'):
    """src/fusionnet_model.py:172-302 canonical branch (single scale, w_smoothness=0):
    GT masked where lidar exists (:214-221), mean |.| over valid GT (:245-248) plus
    w_lidar * mean |.| over valid lidar (:250-253, :293)."""
    fn = {'l1': F.l1_loss, 'l2': F.mse_loss, 'smoothl1': F.smooth_l1_loss}[loss_func]
    if w_lidar_loss > 0.0:
        ground_truth = ground_truth * (lidar_map <= 0.0).to(ground_truth.dtype)
    v_gt = ground_truth > 0
    v_li = lidar_map > 0
    loss = fn(output_depth[v_gt], ground_truth[v_gt])
    if w_lidar_loss > 0.0:
        loss = loss + w_lidar_loss * fn(output_depth[v_li], lidar_map[v_li])
    return loss


def outlier_removal(depth, kernel_size=7, threshold=1.5):
    """src/net_utils.py:591-638."""
    max_value = 10 * torch.max(depth)
    filled = torch.where(depth > 0, depth, torch.full_like(depth, float(max_value)))
    pad = kernel_size // 2
    filled = F.pad(filled, (pad, pad, pad, pad), mode='constant', value=float(max_value))
    min_values = -F.max_pool2d(-filled, kernel_size, 1, 0)
    keep = torch.where(min_values < depth - threshold, torch.zeros_like(depth), torch.ones_like(depth))
    return depth * keep


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam (weight_decay=0, amsgrad=False) as constructed at
    src/fusionnet_main.py:307-312.  In place on params / moments."""
    b1, b2 = betas
    for p_, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        if g is None:
            continue
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p_.addcdiv_(m, denom, value=-lr / bc1)
