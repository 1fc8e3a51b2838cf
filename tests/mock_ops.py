"""
TEST INFRASTRUCTURE ONLY: a torch-CPU emulation of rcfd.ops with the C-ABI's semantics
(include/rcfd.h), used by the ``not gpu`` tests to exercise the HOST logic (engine graph
wiring, reverse-mode tape, gradient delivery, data-parallel sync) in the build container,
which has no GPU.  Never imported by the product.
"""
import sys

import torch
import torch.nn.functional as F

F32, BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_SIGMOID, ACT_DEPTH_HEAD = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TMA, ENGINE_STRIP = 0, 1, 2, 3, 4
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class hold_allocations(object):
    """rcfd.ops.hold_allocations: nothing to hold on one (CPU) stream."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _nchw(x):
    return x.permute(0, 3, 1, 2).float()


def _nhwc(x, dtype=torch.float32):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def conv_out_size(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def _act(v, act, params=(0.0, 0.0)):
    if act == ACT_LEAKY:
        return F.leaky_relu(v, 0.2)
    if act == ACT_SIGMOID:
        return torch.sigmoid(v)
    if act == ACT_DEPTH_HEAD:
        return params[0] / (torch.sigmoid(v) + params[1])
    return v


def _gather_input(x0, x1, in_size, in_dilation):
    a = _nchw(x0)
    if in_size is not None and tuple(in_size) != tuple(a.shape[-2:]):
        a = F.interpolate(a, size=tuple(int(v) for v in in_size))
    if x1 is not None:
        a = torch.cat([a, _nchw(x1)], 1)
    return a


def conv2d(x0, weight_packed, cout, k, stride=1, x1=None, in_size=None, scale=None, shift=None, act=ACT_NONE,
           act_params=(0.0, 0.0), residual=None, stats=None, out=None, out_f32=False, accumulate=False,
           in_dilation=1, out_size=None, pad=None, engine=ENGINE_AUTO, weight_up2x=None):
    if isinstance(weight_packed, tuple):            # tensor-core parity mode: the REAL orchestration over mocked passes
        from rcfd import x3
        return x3.conv2d_x3(sys.modules[__name__], x0, weight_packed, cout, k, stride, x1, in_size, scale, shift, act,
                            act_params, residual, stats, out, accumulate, in_dilation, out_size, pad, engine, weight_up2x)
    pad = k // 2 if pad is None else pad
    w = weight_packed.float().view(cout, k, k, -1).permute(0, 3, 1, 2)
    if in_dilation == 2:
        a = _nchw(x0)
        n, c, h, wd = a.shape
        ho, wo = out_size
        z = torch.zeros(n, c, max(2 * h - 1, ho + k), max(2 * wd - 1, wo + k))
        z[:, :, 0:2 * h - 1:2, 0:2 * wd - 1:2] = a          # zeros inserted between samples
        y = F.conv2d(z, w, None, 1, pad)[:, :, :ho, :wo]
    else:
        a = _gather_input(x0, x1, in_size, 1)
        y = F.conv2d(a, w, None, stride, pad)
        if out_size is not None:                    # even windows (space-to-depth stem, k = 4 / pad 2): first ho x wo outputs
            y = y[:, :, :out_size[0], :out_size[1]]
            assert tuple(y.shape[-2:]) == tuple(out_size), (y.shape, out_size)
    if stats is not None:
        stats[0].add_(y.double().sum(dim=(0, 2, 3)))
        stats[1].add_((y.double() ** 2).sum(dim=(0, 2, 3)))
    if scale is not None:
        y = y * scale[None, :, None, None] + shift[None, :, None, None]
    y = _act(y, act, act_params)
    if residual is not None:
        y = F.leaky_relu(y + _nchw(residual), 0.2)
    res = _nhwc(y, torch.float32 if out_f32 else x0.dtype)
    if out is not None:
        out.copy_(out + res if accumulate else res)
        return out
    return res


def split_bf16(x, parts=2):
    out, r = [], x.float()
    for _ in range(parts):
        out.append(r.to(torch.bfloat16))
        r = r - out[-1].float()
    return tuple(out)


def channel_stats(y, ssum, ssq):
    c = y.shape[-1]
    ssum.add_(y.double().reshape(-1, c).sum(0))
    ssq.add_((y.double() ** 2).reshape(-1, c).sum(0))


def epilogue_f32(y, scale, shift, act, act_params=(0.0, 0.0), residual=None, out=None):
    v = y if scale is None else y * scale + shift
    v = _act(v, act, act_params)
    if residual is not None:
        v = F.leaky_relu(v + residual, 0.2)
    if out is not None:
        out.copy_(v)
        return out
    return v


def conv2d_wgrad(x0, dy, k, stride=1, x1=None, in_size=None, pad=None, engine=ENGINE_AUTO, x3=0, out=None):
    assert out is None          # persistent buffers belong to the multi-stream (CUDA) schedule
    if x3:
        from rcfd import x3 as x3mod
        return x3mod.wgrad_x3(sys.modules[__name__], int(x3), x0, dy, k, stride, x1, in_size, pad, engine)
    return _conv2d_wgrad(x0, dy, k, stride, x1, in_size, pad)


@torch.enable_grad()
def _conv2d_wgrad(x0, dy, k, stride, x1, in_size, pad):
    a = _gather_input(x0, x1, in_size, 1)
    cout = dy.shape[3]
    w = torch.zeros(cout, a.shape[1], k, k, requires_grad=True)   # a already carries any channel padding
    y = F.conv2d(a, w, None, stride, k // 2 if pad is None else pad)[:, :, :dy.shape[1], :dy.shape[2]]
    y.backward(_nchw(dy))
    return w.grad.permute(0, 2, 3, 1).reshape(cout, k * k, -1).contiguous()


def pack_weight(w_oihw, dtype, cin_off=0, cin_cnt=None, dgrad=False, out=None, pad_to=None):
    cout, cin, kh, kw = w_oihw.shape
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    w = w_oihw[:, cin_off:cin_off + cin_cnt]
    if dgrad:
        real = w.shape[1]                     # rows past the real channels (a zero-padded source) are zero (include/rcfd.h)
        res = w.flip(2, 3).permute(1, 2, 3, 0).reshape(real, kh * kw, cout)
        if real < cin_cnt:
            res = torch.cat([res, torch.zeros(cin_cnt - real, kh * kw, cout, dtype=res.dtype)], 0)
    else:
        res = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin_cnt)
    if pad_to is not None and pad_to > res.shape[2]:
        res = F.pad(res, (0, pad_to - res.shape[2]))
    res = res.contiguous().to(dtype)
    if out is not None:
        out.copy_(res)
        return out
    return res


def pack_upconv2x_weight(w_oihw, dtype):
    return torch.zeros(4, w_oihw.shape[0], 4, w_oihw.shape[1], dtype=dtype)   # unused by the mock conv


def unpack_wgrad(dw_packed, grad_oihw, cin_off=0, accumulate=False, cin_cnt=None):
    cout, cin, kh, kw = grad_oihw.shape
    cnt = min(dw_packed.shape[2], cin - cin_off) if cin_cnt is None else cin_cnt
    g = dw_packed[:cout].reshape(cout, kh, kw, dw_packed.shape[2])[..., :cnt].permute(0, 3, 1, 2)
    if accumulate:
        grad_oihw[:, cin_off:cin_off + cnt] += g
    else:
        grad_oihw[:, cin_off:cin_off + cnt] = g


def bn_finalize(ssum, ssq, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, count):
    mean = ssum / count
    var = (ssq / count - mean * mean).clamp_min(0)
    invstd = 1.0 / torch.sqrt(var + BN_EPS)
    scale.copy_((gamma.double() * invstd).float())
    shift.copy_((beta.double() - mean * gamma.double() * invstd).float())
    save_mean.copy_(mean.float())
    save_invstd.copy_(invstd.float())
    if running_mean is not None:
        running_mean.copy_(((1 - BN_MOMENTUM) * running_mean.double() + BN_MOMENTUM * mean).float())
        running_var.copy_(((1 - BN_MOMENTUM) * running_var.double() + BN_MOMENTUM * var * count / max(count - 1, 1)).float())


def bn_fold(gamma, beta, running_mean, running_var, scale, shift):
    s = gamma / torch.sqrt(running_var + BN_EPS)
    scale.copy_(s)
    shift.copy_(beta - running_mean * s)


def bn_act(y, scale, shift, act, residual=None, out=None):
    v = y.float()
    if scale is not None:
        v = v * scale + shift
    v = _act(v, act)
    if residual is not None:
        v = F.leaky_relu(v + residual.float(), 0.2)
    return v.to(y.dtype)


def bn_train_act(y, ssum, ssq, bn_weight, bn_bias, running_mean, running_var, scale, shift, save_mean, save_invstd, act,
                 residual=None):
    bn_finalize(ssum, ssq, bn_weight, bn_bias, running_mean, running_var, scale, shift, save_mean, save_invstd,
                y.numel() // y.shape[-1])
    return bn_act(y, scale, shift, act, residual=residual)


def _dact(pre, dz, act):
    if act == ACT_LEAKY:
        return torch.where(pre > 0, dz, 0.2 * dz)
    if act == ACT_SIGMOID:
        s = torch.sigmoid(pre)
        return dz * s * (1 - s)
    return dz


def bn_act_bwd(dz, y, scale, shift, mean, invstd, act, dgamma, dbeta, sums=None, post_z=None):
    if post_z is not None:
        dzm = leaky_bwd(dz, post_z)
        return bn_act_bwd(dzm, y, scale, shift, mean, invstd, act, dgamma, dbeta, sums=sums), dzm
    c = y.shape[-1]
    yf, dzf = y.float().reshape(-1, c), dz.float().reshape(-1, c)
    d = _dact(yf * scale + shift, dzf, act)
    xhat = (yf - mean) * invstd
    s0, s1 = d.double().sum(0), (d * xhat).double().sum(0)
    m = yf.shape[0]
    dy = scale * (d - (s0 / m).float() - xhat * (s1 / m).float())
    dgamma.copy_(s1.float())
    dbeta.copy_(s0.float())
    return dy.view(y.shape).to(y.dtype)


def gate_fuse(y2c, scale, shift, img):
    c = img.shape[-1]
    v = y2c.float()
    if scale is not None:
        v = v * scale + shift
    return (torch.sigmoid(v[..., :c]) * v[..., c:] + img.float()).to(img.dtype)


def gate_fuse_bwd(dout, y2c, scale, shift):
    c = dout.shape[-1]
    v = y2c.float()
    if scale is not None:
        v = v * scale + shift
    s = torch.sigmoid(v[..., :c])
    g = dout.float()
    return torch.cat([g * v[..., c:] * s * (1 - s), g * s], -1).to(y2c.dtype)


def maxpool3x3s2(x):
    return _nhwc(F.max_pool2d(_nchw(x), 3, 2, 1), x.dtype)


@torch.enable_grad()
def maxpool3x3s2_bwd(x, dout):
    a = _nchw(x).requires_grad_(True)
    F.max_pool2d(a, 3, 2, 1).backward(_nchw(dout))
    return _nhwc(a.grad, x.dtype)


@torch.enable_grad()
def upsample_nearest_bwd(dup, src_hw):
    n, hu, wu, c = dup.shape
    a = torch.zeros(n, c, src_hw[0], src_hw[1], requires_grad=True)
    F.interpolate(a, size=(hu, wu)).backward(_nchw(dup))
    return _nhwc(a.grad, dup.dtype)


def leaky_bwd(dout, out):
    return torch.where(out.float() > 0, dout.float(), 0.2 * dout.float()).to(dout.dtype)


def add_(acc, x):
    acc.add_(x)
    return acc


def nchw_to_nhwc(x, dtype, cpad=None):
    y = _nhwc(x, dtype)
    if cpad is not None and cpad > y.shape[3]:
        y = F.pad(y, (0, cpad - y.shape[3]))
    return y


def nchw_to_s2d(x, dtype, cpad=16):
    n, c, h, w = x.shape
    y = x.view(n, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(n, h // 2, w // 2, 4 * c)
    return F.pad(y, (0, cpad - 4 * c)).contiguous().to(dtype)


def pack_stem_s2d_weight(w_oihw, dtype, cpad=16):
    # include/rcfd.h: w'[ty][tx][(dy*2+dx)*c + ch] = w[2ty+dy-1][2tx+dx-1] (zero outside the 7x7 window)
    cout, c, _, _ = w_oihw.shape
    out = torch.zeros(cout, 4, 4, cpad)
    for ty in range(4):
        for tx in range(4):
            for dy in range(2):
                for dx in range(2):
                    r, q = 2 * ty + dy - 1, 2 * tx + dx - 1
                    if 0 <= r < 7 and 0 <= q < 7:
                        out[:, ty, tx, (dy * 2 + dx) * c:(dy * 2 + dx + 1) * c] = w_oihw[:, :, r, q]
    return out.view(cout, 16, cpad).to(dtype)


def unpack_stem_s2d_wgrad(dw_packed, grad_oihw):
    cout, c, _, _ = grad_oihw.shape
    d = dw_packed.view(-1, 4, 4, dw_packed.shape[2])
    for ty in range(4):
        for tx in range(4):
            for dy in range(2):
                for dx in range(2):
                    r, q = 2 * ty + dy - 1, 2 * tx + dx - 1
                    if 0 <= r < 7 and 0 <= q < 7:
                        grad_oihw[:, :, r, q] = d[:cout, ty, tx, (dy * 2 + dx) * c:(dy * 2 + dx + 1) * c]


def nhwc_to_nchw(x):
    return _nchw(x).contiguous()


def depth_head_bwd(ddepth, depth, min_depth, min_over_max, dtype, cpad=1):
    s = min_depth / depth - min_over_max
    g = (-ddepth.reshape(depth.shape) * depth * depth / min_depth * s * (1 - s)).to(dtype)
    return F.pad(g, (0, cpad - 1)) if cpad > 1 else g


@torch.enable_grad()
def masked_l1_loss(out, gt, lidar, w_lidar, want_grad=True):
    o = out.detach().clone().requires_grad_(True)
    g = gt * (lidar <= 0).float()
    vg, vl = g > 0, lidar > 0
    loss = (o[vg] - g[vg]).abs().mean() + w_lidar * (o[vl] - lidar[vl]).abs().mean()
    loss.backward()
    return loss.detach().view(1), o.grad


def outlier_removal(depth, kernel_size=7, threshold=1.5):
    mx = 10 * depth.max()
    filled = torch.where(depth > 0, depth, mx.expand_as(depth))
    p = kernel_size // 2
    mn = -F.max_pool2d(-F.pad(filled, (p, p, p, p), value=float(mx)), kernel_size, 1, 0)
    return torch.where(mn < depth - threshold, torch.zeros_like(depth), depth)


def bilinear2x(x):
    n, h, w, c = x.shape
    y = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2, mode='bilinear', align_corners=True)
    return y.permute(0, 2, 3, 1).contiguous()


def bilinear2x_bwd(dy):
    n, ho, wo, c = dy.shape
    with torch.enable_grad():                 # called from inside an autograd.Function.backward
        x = torch.zeros(n, 1, ho // 2, wo // 2, requires_grad=True)
        y = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
        (g,) = torch.autograd.grad(y, x, dy.permute(0, 3, 1, 2).float())
    return g.permute(0, 2, 3, 1).contiguous()


def concat_logit(skip, logit, dtype, cpad=16):
    n, h, w, _ = logit.shape
    c = 0 if skip is None else skip.shape[3]
    co = (c + 1 + cpad - 1) // cpad * cpad
    out = torch.zeros(n, h, w, co, dtype=dtype)
    if skip is not None:
        out[..., :c] = skip
    out[..., c] = logit[..., 0].to(dtype)
    return out


def split_logit(dcat, c):
    dskip = dcat[..., :c].contiguous() if c > 0 else None
    return dskip, dcat[..., c:c + 1].float().contiguous()
