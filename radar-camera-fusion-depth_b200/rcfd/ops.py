"""
Tensor-level wrappers over the C-ABI: torch is used for device memory and streams only.
Every function requires CUDA tensors and calls librcfd_b200.so; nothing here computes in
PyTorch.  Activations are NHWC tensors [N, H, W, C] (float32 or bfloat16).
"""
import ctypes
import os
import sys

import torch

from . import _lib
from ._lib import (ConvDesc, F32, BF16, ACT_NONE, ACT_LEAKY, ACT_SIGMOID, ACT_DEPTH_HEAD,
                   ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TMA, ENGINE_STRIP)

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def dt(t):
    return _DT[t.dtype]


# While several CUDA streams are in flight (rcfd.engine multi-stream schedules) every tensor these wrappers
# allocate is kept alive until the streams have joined: torch's caching allocator recycles a freed block on
# the stream that allocated it, which is only safe once every OTHER stream that read it has been joined.
_HOLD = None


class hold_allocations(object):
    """Context manager: keep every ops-allocated tensor alive until exit (re-entrant: the outermost owns)."""

    def __enter__(self):
        global _HOLD
        self.owner = _HOLD is None
        if self.owner:
            _HOLD = []
        return self

    def __exit__(self, *exc):
        global _HOLD
        if self.owner:
            _HOLD = None
        return False


def _keep(t):
    if _HOLD is not None:
        _HOLD.append(t)
    return t


def _empty(*a, **k):
    return _keep(torch.empty(*a, **k))


def _empty_like(*a, **k):
    return _keep(torch.empty_like(*a, **k))


def _zeros(*a, **k):
    return _keep(torch.zeros(*a, **k))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        raise RuntimeError('rcfd ops need contiguous CUDA tensors (there is no CPU fallback)')
    return ctypes.c_void_p(t.data_ptr())


def set_option(key, value):
    _lib.call('rcfd_set_option', key.encode(), int(value))


def conv_out_size(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def split_bf16(x, parts=2):
    """float32 tensor -> ``parts`` (2 or 3) bfloat16 tensors of the same shape whose sum reproduces x to ~2^-17
    (2 parts) / ~2^-25 (3 parts) relative: operand split of the tensor-core parity modes (include/rcfd.h)."""
    x = x.contiguous()
    assert x.dtype == torch.float32 and parts in (2, 3)
    out = tuple(_empty(x.shape, device=x.device, dtype=torch.bfloat16) for _ in range(parts))
    _lib.call('rcfd_split_bf16', _p(x), _p(out[0]), _p(out[1]), _p(out[2]) if parts == 3 else None, x.numel(), _stream())
    return out


def channel_stats(y, ssum, ssq):
    """Per-channel sum / sum of squares of an fp32 NHWC tensor, ADDED to the float64 [C] tensors."""
    c = y.shape[-1]
    _lib.call('rcfd_channel_stats', _p(y), _p(ssum), _p(ssq), y.numel() // c, c, _stream())


def epilogue_f32(y, scale, shift, act, act_params=(0.0, 0.0), residual=None, out=None):
    """out = act(y * scale + shift) [, leaky(out + residual)] on fp32 NHWC tensors (any channel count)."""
    c = y.shape[-1]
    out = _empty_like(y) if out is None else out
    _lib.call('rcfd_epilogue_f32', _p(y), _p(scale), _p(shift), _p(residual), _p(out), y.numel() // c, c, act,
              float(act_params[0]), float(act_params[1]), _stream())
    return out


def conv2d(x0, weight_packed, cout, k, stride=1, x1=None, in_size=None, scale=None, shift=None,
           act=ACT_NONE, act_params=(0.0, 0.0), residual=None, stats=None, out=None, out_f32=False,
           accumulate=False, in_dilation=1, out_size=None, pad=None, engine=ENGINE_AUTO, weight_up2x=None):
    """Implicit-GEMM convolution (see include/rcfd.h rcfd_conv2d_fwd).
    x0: [N, h0, w0, c0]; in_size: logical (H, W) the taps see (x0 is nearest-up-sampled to it);
    x1: optional [N, H, W, c1] concat partner; stats: (sum, sqsum) float64 [cout] tensors.
    weight_packed given as a tuple of 2 / 3 bf16 parts selects the tensor-core parity mode (fp32 operands)."""
    if isinstance(weight_packed, tuple):
        from . import x3
        return x3.conv2d_x3(sys.modules[__name__], x0, weight_packed, cout, k, stride, x1, in_size, scale, shift, act,
                            act_params, residual, stats, out, accumulate, in_dilation, out_size, pad, engine, weight_up2x)
    for t in (x0, x1, weight_packed, scale, shift, residual, out):
        _p(t)                                   # CUDA + contiguity check
    n, h0, w0, c0 = x0.shape
    hin, win = (h0, w0) if in_size is None else (int(in_size[0]), int(in_size[1]))
    pad = k // 2 if pad is None else pad
    if out_size is None:
        if in_dilation == 1:
            ho, wo = conv_out_size(hin, k, stride, pad), conv_out_size(win, k, stride, pad)
        else:
            raise ValueError('out_size is required with in_dilation')
    else:
        ho, wo = out_size
    d = ConvDesc()
    d.n, d.ho, d.wo, d.cout = n, ho, wo, cout
    d.kh = d.kw = k
    d.stride, d.pad, d.in_dilation = stride, pad, in_dilation
    d.hin, d.win = hin, win
    d.src0, d.h0, d.w0, d.c0 = x0.data_ptr(), h0, w0, c0
    if x1 is not None:
        assert x1.shape[0] == n and x1.shape[1] == hin and x1.shape[2] == win and x1.dtype == x0.dtype
        d.src1, d.c1 = x1.data_ptr(), x1.shape[3]
    else:
        d.src1, d.c1 = None, 0
    d.weight = weight_packed.data_ptr()
    if out is None:
        out = _empty((n, ho, wo, cout), device=x0.device, dtype=torch.float32 if out_f32 else x0.dtype)
    d.dst = out.data_ptr()
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.act, d.act_p0, d.act_p1 = act, act_params[0], act_params[1]
    d.residual = residual.data_ptr() if residual is not None else None
    if stats is not None:
        d.stats_sum, d.stats_sqsum = stats[0].data_ptr(), stats[1].data_ptr()
    d.accumulate = 1 if accumulate else 0
    d.dst_f32 = 1 if out_f32 else 0
    d.dtype = dt(x0)
    d.engine = engine
    d.weight_up2x = weight_up2x.data_ptr() if weight_up2x is not None else None
    _lib.call('rcfd_conv2d_fwd', ctypes.byref(d), _stream())
    return out


def conv2d_wgrad(x0, dy, k, stride=1, x1=None, in_size=None, pad=None, engine=ENGINE_AUTO, x3=0, out=None):
    """Packed float weight gradient [cout][k*k][c0+c1] of the convolution above.  x3 = 2 | 3: tensor-core parity
    mode (fp32 operands split into that many bf16 parts; 3 / 6 bf16 passes summed in fp32, rcfd/x3.py).
    out: destination to overwrite (the persistent buffers of the batched unpack)."""
    if x3:
        assert out is None
        from . import x3 as x3mod
        return x3mod.wgrad_x3(sys.modules[__name__], int(x3), x0, dy, k, stride, x1, in_size, pad, engine)
    for t in (x0, x1, dy):
        _p(t)
    n, h0, w0, c0 = x0.shape
    hin, win = (h0, w0) if in_size is None else (int(in_size[0]), int(in_size[1]))
    pad = k // 2 if pad is None else pad
    _, ho, wo, cout = dy.shape
    c1 = 0 if x1 is None else x1.shape[3]
    d = ConvDesc()
    d.n, d.ho, d.wo, d.cout = n, ho, wo, cout
    d.kh = d.kw = k
    d.stride, d.pad, d.in_dilation = stride, pad, 1
    d.hin, d.win = hin, win
    d.src0, d.h0, d.w0, d.c0 = x0.data_ptr(), h0, w0, c0
    d.src1, d.c1 = (x1.data_ptr() if x1 is not None else None), c1
    dw = _empty((cout, k * k, c0 + c1), device=x0.device, dtype=torch.float32) if out is None else out
    assert dw.shape == (cout, k * k, c0 + c1) and dw.dtype == torch.float32 and dw.is_contiguous()
    d.weight = dw.data_ptr()       # unused by wgrad, must be non-null
    d.dst = dy.data_ptr()
    d.dtype = dt(x0)
    d.engine = engine
    ws_bytes = int(_lib.load().rcfd_conv2d_wgrad_workspace(ctypes.byref(d)))
    ws = _empty(ws_bytes, device=x0.device, dtype=torch.uint8) if ws_bytes > 0 else None
    _lib.call('rcfd_conv2d_wgrad', ctypes.byref(d), _p(dw), _p(ws), ws_bytes, _stream())
    return dw


def pack_weight(w_oihw, dtype, cin_off=0, cin_cnt=None, dgrad=False, out=None, pad_to=None):
    """OIHW float -> packed [cout][taps][cin_pad] (forward) or [cin_cnt][taps][cout_pad] (dgrad);
    pad_to zero-pads the innermost (channel) dimension."""
    cout, cin, kh, kw = w_oihw.shape
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    inner = cout if dgrad else cin_cnt
    pad = inner if pad_to is None else max(inner, int(pad_to))
    if out is None:
        shape = (cin_cnt, kh * kw, pad) if dgrad else (cout, kh * kw, pad)
        out = _empty(shape, device=w_oihw.device, dtype=dtype)
    _lib.call('rcfd_pack_conv_weight', _p(w_oihw), _p(out), cout, cin, kh, kw, cin_off, cin_cnt, pad,
              1 if dgrad else 0, _DT[dtype], _stream())
    return out


def pack_upconv2x_weight(w_oihw, dtype):
    """[4 phases][cout][2x2 taps][cin] weights of `3x3 conv after 2x nearest up-sampling` (TMA engine)."""
    cout, cin, kh, kw = w_oihw.shape
    assert kh == 3 and kw == 3
    out = _empty((4, cout, 4, cin), device=w_oihw.device, dtype=dtype)
    _lib.call('rcfd_pack_upconv2x_weight', _p(w_oihw), _p(out), cout, cin, _DT[dtype], _stream())
    return out


def pack_dgrad_s2_weight(w_oihw, dtype, cin_off=0, cin_cnt=None, pad_to=None):
    """[4 phases][cin_cnt][2x2 taps][cout_pad] phase weights of the data gradient of a 3x3 / stride-2 conv (include/rcfd.h)."""
    cout, cin, kh, kw = w_oihw.shape
    assert kh == 3 and kw == 3
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    pad = cout if pad_to is None else max(cout, int(pad_to))
    out = _empty((4, cin_cnt, 4, pad), device=w_oihw.device, dtype=dtype)
    _lib.call('rcfd_pack_dgrad_s2_weight', _p(w_oihw), _p(out), cout, cin, cin_off, cin_cnt, pad, _DT[dtype], _stream())
    return out


def pack_upconv2x_dgrad_weight(w_oihw, dtype, cin_off=0, cin_cnt=None, pad_to=None):
    """[cin_cnt][4x4 taps][cout_pad]: forward-packed weights of the 4x4 / stride-2 conv over dy that is the data gradient of a
    3x3 conv behind an exact 2x nearest up-sampling w.r.t. its low-res source (include/rcfd.h)."""
    cout, cin, kh, kw = w_oihw.shape
    assert kh == 3 and kw == 3
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    pad = cout if pad_to is None else max(cout, int(pad_to))
    out = _empty((cin_cnt, 16, pad), device=w_oihw.device, dtype=dtype)
    _lib.call('rcfd_pack_upconv2x_dgrad_weight', _p(w_oihw), _p(out), cout, cin, cin_off, cin_cnt, pad, _DT[dtype], _stream())
    return out


def unpack_wgrad(dw_packed, grad_oihw, cin_off=0, accumulate=False, cin_cnt=None):
    """packed float [>=cout][taps][cin_pad] -> OIHW slice; cin_cnt defaults to the packed width."""
    cout, cin, kh, kw = grad_oihw.shape
    cin_pad = dw_packed.shape[2]
    cin_cnt = min(cin_pad, cin - cin_off) if cin_cnt is None else cin_cnt
    _lib.call('rcfd_unpack_conv_wgrad', _p(dw_packed), _p(grad_oihw), cout, cin, kh, kw, cin_off, cin_cnt, cin_pad,
              1 if accumulate else 0, _stream())


# ---------------------------------------------------------------------------------- batched (un)packing
PACK_FWD, PACK_DGRAD, PACK_UP2X, PACK_STEM_S2D, UNPACK_CONV, UNPACK_STEM_S2D, COPY_F32, PACK_DGRAD_S2, PACK_UPCONV_DGRAD = range(9)


def spec_pack_weight(w, dtype, cin_off=0, cin_cnt=None, dgrad=False, pad_to=None):
    """What pack_weight(w, ...) would do, as data: (shape, dtype, needs_zero_init, [item fields]) for PackBatch."""
    cout, cin, kh, kw = w.shape
    taps = kh * kw
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    inner = cout if dgrad else cin_cnt
    pad = inner if pad_to is None else max(inner, int(pad_to))
    if dgrad:
        item = dict(kind=PACK_DGRAD, src=w, off=0, total=cin_cnt * taps * cout, cout=cout, cin=cin, taps=taps,
                    cin_off=cin_off, cin_cnt=cin_cnt, cpad=pad, col_off=0, dst_cols=pad)
        return (cin_cnt, taps, pad), dtype, pad > cout, [item]
    item = dict(kind=PACK_FWD, src=w, off=0, total=cout * taps * pad, cout=cout, cin=cin, taps=taps, cin_off=cin_off,
                cin_cnt=cin_cnt, cpad=pad)
    return (cout, taps, pad), dtype, False, [item]


def spec_pack_stacked_1x1(ws, dtype, dgrad=False):
    """pack_weight(torch.cat(ws, 0), dtype, dgrad=dgrad) for 1x1 weights without the cat: one item per weight."""
    cin = ws[0].shape[1]
    ctot = sum(w.shape[0] for w in ws)
    items, row = [], 0
    for w in ws:
        c = w.shape[0]
        assert w.shape[1] == cin and w.shape[2] == 1 and w.shape[3] == 1
        if dgrad:
            items.append(dict(kind=PACK_DGRAD, src=w, off=0, total=cin * c, cout=c, cin=cin, taps=1, cin_off=0,
                              cin_cnt=cin, cpad=ctot, col_off=row, dst_cols=ctot))
        else:
            items.append(dict(kind=PACK_FWD, src=w, off=row * cin, total=c * cin, cout=c, cin=cin, taps=1, cin_off=0,
                              cin_cnt=cin, cpad=cin))
        row += c
    return ((cin, 1, ctot) if dgrad else (ctot, 1, cin)), dtype, False, items


def spec_pack_upconv2x_weight(w, dtype):
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return (4, cout, 4, cin), dtype, False, [dict(kind=PACK_UP2X, src=w, off=0, total=16 * cout * cin, cout=cout, cin=cin,
                                                  taps=9)]


def spec_pack_dgrad_s2_weight(w, dtype, cin_off=0, cin_cnt=None, pad_to=None):
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    pad = cout if pad_to is None else max(cout, int(pad_to))
    return (4, cin_cnt, 4, pad), dtype, False, [dict(kind=PACK_DGRAD_S2, src=w, off=0, total=16 * cin_cnt * pad, cout=cout,
                                                     cin=cin, taps=9, cin_off=cin_off, cin_cnt=cin_cnt, cpad=pad)]


def spec_pack_upconv2x_dgrad_weight(w, dtype, cin_off=0, cin_cnt=None, pad_to=None):
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    cin_cnt = cin - cin_off if cin_cnt is None else cin_cnt
    pad = cout if pad_to is None else max(cout, int(pad_to))
    return (cin_cnt, 16, pad), dtype, False, [dict(kind=PACK_UPCONV_DGRAD, src=w, off=0, total=cin_cnt * 16 * pad, cout=cout,
                                                   cin=cin, taps=9, cin_off=cin_off, cin_cnt=cin_cnt, cpad=pad)]


def spec_pack_stem_s2d_weight(w, dtype, cpad=16):
    cout, c, kh, kw = w.shape
    assert kh == 7 and kw == 7 and 4 * c <= cpad
    return (cout, 16, cpad), dtype, False, [dict(kind=PACK_STEM_S2D, src=w, off=0, total=cout * 16 * cpad, cout=cout,
                                                 cin=c, taps=16, cpad=cpad)]


class PackBatch(object):
    """Device table of rcfd_pack_item for rcfd_pack_batch (include/rcfd.h): ONE launch packs every weight of a training
    step, another one unpacks every weight gradient.  add() takes the fields of one item; src / dst are tensors (kept
    alive here), off / src_off = element offsets into dst / src."""

    def __init__(self):
        self.rows = []
        self.keep = []
        self.table = None
        self.total_blocks = 0

    def add(self, kind, src, dst, off=0, total=0, cout=0, cin=0, taps=0, cin_off=0, cin_cnt=0, cpad=0, col_off=0,
            dst_cols=0, src_off=0):
        assert self.table is None and total > 0
        for t in (src, dst):
            _p(t)
        it = _lib.PackItem(src.data_ptr() + src_off * src.element_size(), dst.data_ptr() + off * dst.element_size(), total,
                           kind, _DT[dst.dtype], cout, cin, taps, cin_off, cin_cnt, cpad, col_off, dst_cols,
                           self.total_blocks, 0)
        it.nblocks = int(_lib.load().rcfd_pack_item_blocks(ctypes.byref(it)))
        if it.nblocks <= 0:
            raise _lib.RcfdError('rcfd_pack_item_blocks: invalid item')
        self.rows.append(it)
        self.total_blocks += it.nblocks
        self.keep += [src, dst]

    def finalize(self, device):
        import numpy as np
        assert ctypes.sizeof(_lib.PackItem) == 72           # sizeof(rcfd_pack_item)
        arr = (_lib.PackItem * len(self.rows))(*self.rows)
        host = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.table = torch.from_numpy(host).to(device)
        owner = np.repeat(np.arange(len(self.rows), dtype=np.int32), [it.nblocks for it in self.rows])
        self.block_item = torch.from_numpy(owner).to(device)
        return self

    def run(self):
        _lib.call('rcfd_pack_batch', _p(self.table), _p(self.block_item), len(self.rows), self.total_blocks, _stream())


def bn_finalize(ssum, ssq, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, count):
    _lib.call('rcfd_bn_finalize', _p(ssum), _p(ssq), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
              _p(scale), _p(shift), _p(save_mean), _p(save_invstd), gamma.numel(), int(count), BN_EPS, BN_MOMENTUM,
              _stream())


def bn_fold(gamma, beta, running_mean, running_var, scale, shift):
    _lib.call('rcfd_bn_fold', _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(scale), _p(shift),
              gamma.numel(), BN_EPS, _stream())


def bn_act(y, scale, shift, act, residual=None, out=None):
    c = y.shape[-1]
    out = _empty_like(y) if out is None else out
    _lib.call('rcfd_bn_act_fwd', _p(y), _p(scale), _p(shift), _p(residual), _p(out), y.numel() // c, c, act, dt(y),
              _stream())
    return out


def bn_train_act(y, ssum, ssq, bn_weight, bn_bias, running_mean, running_var, scale, shift, save_mean, save_invstd, act,
                 residual=None):
    """Training-mode BatchNorm + activation (+ residual add + leaky) in one pass: batch statistics -> scale / shift,
    running-statistics update, normalise.  scale / shift / save_mean / save_invstd are outputs kept for backward."""
    c = y.shape[-1]
    out = _empty_like(y)
    _lib.call('rcfd_bn_train_act_fwd', _p(y), _p(ssum), _p(ssq), _p(bn_weight), _p(bn_bias), _p(running_mean),
              _p(running_var), _p(scale), _p(shift), _p(save_mean), _p(save_invstd), _p(residual), _p(out),
              y.numel() // c, c, act, BN_EPS, BN_MOMENTUM, dt(y), _stream())
    return out


# BatchNorm backward of maps up to this many pixels (N * H * W) runs as ONE launch (rcfd_bn_act_bwd_fused); 0 = never
BN_BWD_FUSED_MAX_PIXELS = int(os.environ.get('RCFD_BN_BWD_FUSED_MAX_PIXELS', '1024'))


def bn_act_bwd(dz, y, scale, shift, mean, invstd, act, dgamma, dbeta, sums=None, post_z=None):
    """sums: optional ZEROED float64 [2c] scratch (a slice of a pool zeroed once per step); default: a fresh buffer
    zeroed by the call.  post_z: dz is the gradient of the output of a LeakyReLU applied AFTER a residual add (post_z = that
    output): it is masked first and (dy, masked dz) is returned -- the masked gradient also belongs to the shortcut."""
    c = y.shape[-1]
    pixels = y.numel() // c
    if pixels <= BN_BWD_FUSED_MAX_PIXELS and c % (8 if y.dtype == torch.bfloat16 else 4) == 0:
        dy = _empty_like(y)
        dzm = _empty_like(dz) if post_z is not None else None
        _lib.call('rcfd_bn_act_bwd_fused', _p(dz), _p(y), _p(post_z), _p(dzm), _p(scale), _p(shift), _p(mean), _p(invstd),
                  _p(dy), _p(dgamma), _p(dbeta), pixels, c, act, dt(y), _stream())
        return dy if post_z is None else (dy, dzm)
    if post_z is not None:
        dz = leaky_bwd(dz, post_z)
    entry = 'rcfd_bn_act_bwd_reduce_acc'
    if sums is None:
        sums = _empty(2 * c, device=y.device, dtype=torch.float64)
        entry = 'rcfd_bn_act_bwd_reduce'
    _lib.call(entry, _p(dz), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd), _p(sums), pixels, c,
              act, dt(y), _stream())
    dy = _empty_like(y)
    _lib.call('rcfd_bn_act_bwd_apply', _p(dz), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd), _p(sums), _p(dy),
              _p(dgamma), _p(dbeta), pixels, c, act, dt(y), _stream())
    return dy if post_z is None else (dy, dz)


def gate_fuse(y2c, scale, shift, img):
    c = img.shape[-1]
    out = _empty_like(img)
    _lib.call('rcfd_gate_fuse_fwd', _p(y2c), _p(scale), _p(shift), _p(img), _p(out), img.numel() // c, c, dt(img),
              _stream())
    return out


def gate_fuse_bwd(dout, y2c, scale, shift):
    c = dout.shape[-1]
    dz = _empty_like(y2c)
    _lib.call('rcfd_gate_fuse_bwd', _p(dout), _p(y2c), _p(scale), _p(shift), _p(dz), dout.numel() // c, c, dt(dout),
              _stream())
    return dz


def maxpool3x3s2(x):
    n, h, w, c = x.shape
    out = _empty((n, conv_out_size(h, 3, 2, 1), conv_out_size(w, 3, 2, 1), c), device=x.device, dtype=x.dtype)
    _lib.call('rcfd_maxpool3x3s2_fwd', _p(x), _p(out), n, h, w, c, dt(x), _stream())
    return out


def maxpool3x3s2_bwd(x, dout):
    n, h, w, c = x.shape
    dx = _empty_like(x)
    _lib.call('rcfd_maxpool3x3s2_bwd', _p(x), _p(dout), _p(dx), n, h, w, c, dt(x), _stream())
    return dx


def maxpool3x3s2_idx(x):
    """Training forward: (pooled, positions) -- positions feed maxpool3x3s2_bwd_idx."""
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, 3, 2, 1), conv_out_size(w, 3, 2, 1)
    out = _empty((n, ho, wo, c), device=x.device, dtype=x.dtype)
    idx = _empty((n, ho, wo, c), device=x.device, dtype=torch.uint8)
    _lib.call('rcfd_maxpool3x3s2_fwd_idx', _p(x), _p(out), _p(idx), n, h, w, c, dt(x), _stream())
    return out, idx


def maxpool3x3s2_bwd_idx(dout, idx, in_hw):
    n, ho, wo, c = dout.shape
    h, w = in_hw
    dx = _empty((n, h, w, c), device=dout.device, dtype=dout.dtype)
    _lib.call('rcfd_maxpool3x3s2_bwd_idx', _p(dout), _p(idx), _p(dx), n, h, w, c, dt(dout), _stream())
    return dx


def upsample_nearest_bwd(dup, src_hw):
    n, hu, wu, c = dup.shape
    hs, ws = src_hw
    dsrc = _empty((n, hs, ws, c), device=dup.device, dtype=dup.dtype)
    _lib.call('rcfd_upsample_nearest_bwd', _p(dup), _p(dsrc), n, hs, ws, hu, wu, c, 0, dt(dup), _stream())
    return dsrc


def leaky_bwd(dout, out):
    din = _empty_like(dout)
    _lib.call('rcfd_leaky_bwd', _p(dout), _p(out), _p(din), dout.numel(), dt(dout), _stream())
    return din


def add_(acc, x):
    _lib.call('rcfd_add_inplace', _p(acc), _p(x), acc.numel(), dt(acc), _stream())
    return acc


def nchw_to_nhwc(x, dtype, cpad=None):
    n, c, h, w = x.shape
    cpad = c if cpad is None else max(c, int(cpad))
    x = x.contiguous()
    out = _empty((n, h, w, cpad), device=x.device, dtype=dtype)
    _lib.call('rcfd_nchw_to_nhwc', _p(x), _p(out), n, c, h, w, cpad, _DT[dtype], _stream())
    return out


def nchw_to_s2d(x, dtype, cpad=16):
    """NCHW float -> space-to-depth NHWC [N, H/2, W/2, cpad] (channel = (dy*2+dx)*C + c), for the 7x7/s2 stems."""
    n, c, h, w = x.shape
    x = x.contiguous()
    out = _empty((n, h // 2, w // 2, cpad), device=x.device, dtype=dtype)
    _lib.call('rcfd_nchw_to_s2d_nhwc', _p(x), _p(out), n, c, h, w, cpad, _DT[dtype], _stream())
    return out


def pack_stem_s2d_weight(w_oihw, dtype, cpad=16):
    cout, c, kh, kw = w_oihw.shape
    assert kh == 7 and kw == 7
    out = _empty((cout, 16, cpad), device=w_oihw.device, dtype=dtype)
    _lib.call('rcfd_pack_stem_s2d_weight', _p(w_oihw), _p(out), cout, c, cpad, _DT[dtype], _stream())
    return out


def unpack_stem_s2d_wgrad(dw_packed, grad_oihw):
    cout, c, _, _ = grad_oihw.shape
    _lib.call('rcfd_unpack_stem_s2d_wgrad', _p(dw_packed), _p(grad_oihw), cout, c, dw_packed.shape[2], _stream())


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    out = _empty((n, c, h, w), device=x.device, dtype=torch.float32)
    _lib.call('rcfd_nhwc_to_nchw', _p(x), _p(out), n, c, h, w, dt(x), _stream())
    return out


def depth_head_bwd(ddepth, depth, min_depth, min_over_max, dtype, cpad=1):
    """depth: [N, H, W, 1] float -> dlogit [N, H, W, cpad] (channel 0 carries the gradient)."""
    dl = _empty(tuple(depth.shape[:3]) + (cpad,), device=depth.device, dtype=dtype)
    _lib.call('rcfd_depth_head_bwd', _p(ddepth.contiguous()), _p(depth), _p(dl), float(min_depth), float(min_over_max),
              depth.numel(), cpad, _DT[dtype], _stream())
    return dl


def masked_l1_loss(out, gt, lidar, w_lidar, want_grad=True):
    accum = _empty(4, device=out.device, dtype=torch.float64)
    loss = _empty(1, device=out.device, dtype=torch.float32)
    dout = _empty_like(out) if want_grad else None
    _lib.call('rcfd_masked_l1_loss', _p(out.contiguous()), _p(gt.contiguous()), _p(lidar.contiguous()), float(w_lidar),
              _p(accum), _p(loss), _p(dout), out.numel(), _stream())
    return loss, dout


def bilinear2x(x):
    """float [N, H, W, 1] -> [N, 2H, 2W, 1]: bilinear, align_corners=True (multi-resolution decoder)."""
    n, h, w, c = x.shape
    assert c == 1 and x.dtype == torch.float32
    y = _empty((n, 2 * h, 2 * w, 1), device=x.device, dtype=torch.float32)
    _lib.call('rcfd_bilinear2x_fwd', _p(x), _p(y), n, h, w, _stream())
    return y


def bilinear2x_bwd(dy):
    n, ho, wo, c = dy.shape
    assert c == 1 and dy.dtype == torch.float32 and ho % 2 == 0 and wo % 2 == 0
    dx = _empty((n, ho // 2, wo // 2, 1), device=dy.device, dtype=torch.float32)
    _lib.call('rcfd_bilinear2x_bwd', _p(dy), _p(dx), n, ho // 2, wo // 2, _stream())
    return dx


def concat_logit(skip, logit, dtype, cpad=16):
    """NHWC [skip | logit | zeros] with (C + 1) rounded up to a multiple of cpad channels; skip may be None."""
    n, h, w, _ = logit.shape
    c = 0 if skip is None else skip.shape[3]
    co = (c + 1 + cpad - 1) // cpad * cpad
    out = _empty((n, h, w, co), device=logit.device, dtype=dtype)
    _lib.call('rcfd_concat_logit', _p(skip), _p(logit), _p(out), n * h * w, c, co, _DT[dtype], _stream())
    return out


def split_logit(dcat, c):
    """Transpose of concat_logit: (d skip [N, H, W, c] in dcat's dtype or None, d logit float [N, H, W, 1])."""
    n, h, w, co = dcat.shape
    dskip = _empty((n, h, w, c), device=dcat.device, dtype=dcat.dtype) if c > 0 else None
    dlogit = _empty((n, h, w, 1), device=dcat.device, dtype=torch.float32)
    _lib.call('rcfd_split_logit', _p(dcat), _p(dskip), _p(dlogit), n * h * w, c, co, dt(dcat), _stream())
    return dskip, dlogit


def smoothness_loss(predict, image, want_grad=True):
    """(loss [1], d loss / d predict or None) of the reference's smoothness_loss_func; float N x 1 x H x W / N x C x H x W."""
    n, c, h, w = image.shape
    assert predict.shape == (n, 1, h, w) and predict.dtype == torch.float32 and image.dtype == torch.float32
    accum = _empty(2, device=predict.device, dtype=torch.float64)
    loss = _empty(1, device=predict.device, dtype=torch.float32)
    grad = _empty_like(predict) if want_grad else None
    _lib.call('rcfd_smoothness_loss', _p(predict), _p(image), n, c, h, w, _p(accum), _p(loss), _p(grad), _stream())
    return loss, grad


def sobel_smoothness_loss(predict, image, weights, kh, kw, want_grad=True):
    """(loss [1], gradient or None) of the reference's sobel_smoothness_loss_func (kh x kw generalised Sobel filter)."""
    n, c, h, w = image.shape
    assert c == 3 and predict.shape == (n, 1, h, w) and weights.shape == (n, 1, h, w)
    accum = _empty(2, device=predict.device, dtype=torch.float64)
    loss = _empty(1, device=predict.device, dtype=torch.float32)
    grad = _empty_like(predict) if want_grad else None
    scratch = _empty(2 * predict.numel(), device=predict.device, dtype=torch.float32) if want_grad else None
    _lib.call('rcfd_sobel_smoothness_loss', _p(predict), _p(image), _p(weights), n, h, w, int(kh), int(kw), _p(accum), _p(scratch),
              _p(loss), _p(grad), _stream())
    return loss, grad


def outlier_removal(depth, kernel_size=7, threshold=1.5):
    depth = depth.contiguous()
    n, _, h, w = depth.shape
    out = _empty_like(depth)
    scratch = _empty(1, device=depth.device, dtype=torch.float32)
    _lib.call('rcfd_outlier_removal', _p(depth), _p(out), _p(scratch), n, h, w, kernel_size, float(threshold), _stream())
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step):
    _lib.call('rcfd_adam_step', _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr), float(beta1),
              float(beta2), float(eps), int(step), _stream())


def scatter_points_to_depth_map(points_xy, depth, h, w, img=None):
    """S1.  points_xy: [2, N] float64 CUDA, depth: [N] float64.  img given -> z-buffer merge into it."""
    merge = img is not None
    if img is None:
        img = _empty((h, w), device=points_xy.device, dtype=torch.float64)
    npts = points_xy.shape[1]
    _lib.call('rcfd_scatter_points_to_depth_map', _p(points_xy.contiguous()), _p(depth.contiguous()), npts, _p(img), h,
              w, 1 if merge else 0, _stream())
    return img


def scatter_tiles_argmax(crops, points, h, w, compat=True):
    """S2.  crops: [K, 1, ph, pw] float32; points: [K, 3] float32 (x already shifted by +pad)."""
    k, _, ph, pw = crops.shape
    response = _empty((1, h, w), device=crops.device, dtype=torch.float32)
    if compat:
        depth = _empty((1, h, w), device=crops.device, dtype=torch.int64)
        _lib.call('rcfd_scatter_tiles_argmax', _p(crops.contiguous()), _p(points.contiguous()), k, ph, pw, h, w, 1,
                  _p(depth), None, _p(response), _stream())
    else:
        depth = _empty((1, h, w), device=crops.device, dtype=torch.float32)
        _lib.call('rcfd_scatter_tiles_argmax', _p(crops.contiguous()), _p(points.contiguous()), k, ph, pw, h, w, 0,
                  None, _p(depth), _p(response), _stream())
    return depth, response


def stage1_to_stage2(depth, response, quantize_png16=True):
    """(1 x H x W int64 or float32 depth, 1 x H x W float32 response) -> 1 x 2 x H x W float32 FusionNet input_depth."""
    h, w = response.shape[-2:]
    out = _empty((1, 2, h, w), device=response.device, dtype=torch.float32)
    is_i64 = depth.dtype == torch.int64
    if not is_i64:
        depth = depth.float()
    _lib.call('rcfd_stage1_to_stage2', _p(depth.contiguous()) if is_i64 else None, None if is_i64 else _p(depth.contiguous()),
              _p(response.contiguous().float()), _p(out), h, w, 1 if quantize_png16 else 0, _stream())
    return out


def roi_pool(feat, boxes5, out_size, spatial_scale):
    """feat: NHWC; boxes5: [nbox, 5] float32 (batch_index, x1, y1, x2, y2)."""
    n, h, w, c = feat.shape
    nbox = boxes5.shape[0]
    out = _empty((nbox, out_size[0], out_size[1], c), device=feat.device, dtype=feat.dtype)
    _lib.call('rcfd_roi_pool_fwd', _p(feat), _p(boxes5.contiguous()), _p(out), n, h, w, c, nbox, out_size[0],
              out_size[1], float(spatial_scale), dt(feat), _stream())
    return out


def linear_leaky(x, w, b):
    rows, fin = x.shape
    fout = w.shape[0]
    out = _empty((rows, fout), device=x.device, dtype=torch.float32)
    _lib.call('rcfd_linear_leaky_fwd', _p(x.contiguous()), _p(w), _p(b), _p(out), rows, fin, fout, _stream())
    return out


def transform_batch(image, range_maps, params, norm_mode):
    """Batched augmentation (include/rcfd.h rcfd_transform_batch).  image: N x 3 x H x W float or None; range_maps: list
    of N x c x H x W float tensors; params: N x 11 float (see the header).  Returns (image_out, [maps_out])."""
    ref = image if image is not None else range_maps[0]
    n, _, h, w = ref.shape
    image = image.contiguous() if image is not None else None
    range_maps = [m.contiguous() for m in range_maps]
    for t in [image] + range_maps:
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError('transform_batch takes float32 tensors')
    image_out = _empty_like(image) if image is not None else None
    maps_out = [_empty_like(m) for m in range_maps]
    scratch_max = _empty(1, device=ref.device, dtype=torch.int32)
    scratch_sums = _empty(n, device=ref.device, dtype=torch.float64)
    params = params.contiguous()
    for first in range(0, max(len(range_maps), 1), 4):
        chunk, chunk_out = range_maps[first:first + 4], maps_out[first:first + 4]
        k = len(chunk)
        src = (ctypes.c_void_p * 4)(*([m.data_ptr() for m in chunk] + [None] * (4 - k)))
        dst = (ctypes.c_void_p * 4)(*([m.data_ptr() for m in chunk_out] + [None] * (4 - k)))
        chs = (ctypes.c_int32 * 4)(*([m.shape[1] for m in chunk] + [0] * (4 - k)))
        img = image if first == 0 else None
        _lib.call('rcfd_transform_batch', _p(img), _p(image_out) if img is not None else None, src, dst, chs, k, _p(params),
                  _p(scratch_max), _p(scratch_sums), n, h, w, int(norm_mode), _stream())
    return image_out, maps_out


def roi_pool_bwd(feat, boxes5, dout, out_size, spatial_scale):
    """Gradient of roi_pool w.r.t. ``feat`` (NHWC, storage dtype of feat); dout: [nbox, ph, pw, c]."""
    n, h, w, c = feat.shape
    acc = _empty((n, h, w, c), device=feat.device, dtype=torch.float32)
    _lib.call('rcfd_roi_pool_bwd', _p(feat), _p(boxes5.contiguous()), _p(dout.contiguous()), _p(acc), n, h, w, c,
              boxes5.shape[0], out_size[0], out_size[1], float(spatial_scale), dt(feat), _stream())
    if feat.dtype == torch.float32:
        return acc
    out = _empty((n, h, w, c), device=feat.device, dtype=feat.dtype)
    _lib.call('rcfd_cast_f32', _p(acc), _p(out), acc.numel(), dt(feat), _stream())
    return out


def linear_leaky_bwd(x, w, y, dy, want_dx=True):
    """Backward of y = leaky(x w^T + b): returns (dx or None, dw, db), all float32."""
    rows, fin = x.shape
    fout = w.shape[0]
    dpre = _empty((rows, fout), device=x.device, dtype=torch.float32)
    dx = _empty((rows, fin), device=x.device, dtype=torch.float32) if want_dx else None
    dw = _empty((fout, fin), device=x.device, dtype=torch.float32)
    db = _empty((fout,), device=x.device, dtype=torch.float32)
    _lib.call('rcfd_linear_leaky_bwd', _p(x.contiguous()), _p(w), _p(y.contiguous()), _p(dy.contiguous()), _p(dpre), _p(dx),
              _p(dw), _p(db), rows, fin, fout, _stream())
    return dx, dw, db


def bce_logits_loss(logits, target, validity, pos_weight, want_grad=True):
    """sum(v * BCEWithLogits(x, t, pos_weight)) / sum(v) and its gradient w.r.t. the logits (sync-free)."""
    accum = _empty(2, device=logits.device, dtype=torch.float64)
    loss = _empty(1, device=logits.device, dtype=torch.float32)
    dl = _empty_like(logits) if want_grad else None
    _lib.call('rcfd_bce_logits_loss', _p(logits.contiguous()), _p(target.contiguous()), _p(validity.contiguous()),
              float(pos_weight), _p(accum), _p(loss), _p(dl), logits.numel(), _stream())
    return loss, dl


def decode_crop(src, multiplier, out_hw=None, crop_yx=None, out=None, out_channel=0):
    """On-disk samples -> network input (include/rcfd.h rcfd_decode_crop).  src: uint8 [N, H, W, C] (image as decoded
    from the file) or uint16 / int16-viewed [N, H, W] raster of a 16-bit PNG map; returns float32 [N, C, oh, ow] =
    max(src / multiplier, 0) cropped at crop_yx ([N, 2] int32 (y0, x0) on the device) or at the origin.  With ``out``
    ([N, C_total, oh, ow]) the result is written into channels [out_channel, out_channel + C) of it."""
    if src.dtype == torch.uint8:
        bits, (n, sh, sw, c) = 8, src.shape
    elif src.dtype in (torch.uint16, torch.int16):
        bits, (n, sh, sw), c = 16, src.shape, 1
    else:
        raise RuntimeError('decode_crop takes uint8 images or uint16 maps')
    oh, ow = (sh, sw) if out_hw is None else out_hw
    if out is None:
        out = _empty((n, c, oh, ow), device=src.device, dtype=torch.float32)
        out_channel = 0
    assert out.shape[0] == n and out.shape[2:] == (oh, ow) and out_channel + c <= out.shape[1] and out.is_contiguous()
    dst = ctypes.c_void_p(out.data_ptr() + 4 * out_channel * oh * ow)
    _lib.call('rcfd_decode_crop', _p(src.contiguous()), bits, dst, _p(crop_yx), n, sh, sw, c, oh, ow, float(multiplier),
              out.shape[1] * oh * ow, _stream())
    return out


def encode_u16(x, multiplier):
    """float32 tensor -> uint16 samples of the reference's 16-bit PNG files: uint32(x * multiplier) & 0xffff."""
    out = _empty(x.shape, device=x.device, dtype=torch.uint16)
    _lib.call('rcfd_encode_u16', _p(x.contiguous().float()), _p(out), float(multiplier), x.numel(), _stream())
    return out
