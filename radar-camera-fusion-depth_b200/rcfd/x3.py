"""
Tensor-core PARITY modes: orchestration of one convolution / weight gradient as several passes of the bf16
tcgen05 engines over bf16 splits of the fp32 operands (include/rcfd.h, csrc/parity.cu):

    x = x0 + x1 + x2,   x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)
    'bf16x3' (2 parts):  x.w ~= x0.w0 + x1.w0 + x0.w1                            ~2^-16 per product
    'bf16x6' (3 parts):  ... + x1.w1 + x2.w0 + x0.w2                             ~2^-23 per product (fp32-class)

The passes are ordinary ``ops.conv2d`` / ``ops.conv2d_wgrad`` calls on bf16 tensors -- the engine dispatch sees
the same shapes as in the bf16 fast mode, so the SAME kernels (row-streaming, per-tap TMA, sub-pixel up-conv,
gather, TMA / strip / gather wgrad) run -- with the fp32 destination accumulated across passes, smallest terms
first.  What the fast path fuses into the conv epilogue (BatchNorm statistics, folded BN + activation + residual,
depth head) runs afterwards on the fp32 sum.  The reference convolves in fp32 (src/net_utils.py:63-69, 85).

``ops`` is passed in (rcfd.ops, or the torch-CPU emulation the host-logic tests use).
"""

ACT_NONE = 0


def terms(parts):
    """(activation part, weight part) pairs kept, smallest products first."""
    t = [(i, j) for i in range(parts) for j in range(parts) if i + j < parts]
    return sorted(t, key=lambda ij: -(ij[0] + ij[1]))


def conv2d_x3(ops, x0, wparts, cout, k, stride, x1, in_size, scale, shift, act, act_params, residual, stats, out,
              accumulate, in_dilation, out_size, pad, engine, weight_up2x):
    if accumulate:
        raise ValueError('parity mode: the destination is the accumulator of the passes')
    parts = len(wparts)
    x0p = ops.split_bf16(x0, parts)
    x1p = ops.split_bf16(x1, parts) if x1 is not None else (None,) * parts
    up = weight_up2x if weight_up2x is not None else (None,) * parts
    kw = dict(in_size=in_size, in_dilation=in_dilation, out_size=out_size, pad=pad, engine=engine, out_f32=True)
    y = None
    for i, j in terms(parts):
        y = ops.conv2d(x0p[i], wparts[j], cout, k, stride, x1=x1p[i], weight_up2x=up[j], out=y,
                       accumulate=y is not None, **kw)
    if stats is not None:
        ops.channel_stats(y, stats[0], stats[1])
    if scale is not None or act != ACT_NONE or residual is not None or out is not None:
        y = ops.epilogue_f32(y, scale, shift, act, act_params, residual=residual, out=out)
    return y


def wgrad_x3(ops, parts, x0, dy, k, stride, x1, in_size, pad, engine):
    x0p = ops.split_bf16(x0, parts)
    x1p = ops.split_bf16(x1, parts) if x1 is not None else (None,) * parts
    dyp = ops.split_bf16(dy, parts)
    kw = dict(in_size=in_size, pad=pad, engine=engine)
    dw = None
    for i, j in terms(parts):
        t = ops.conv2d_wgrad(x0p[i], dyp[j], k, stride, x1=x1p[i], **kw)
        dw = t if dw is None else ops.add_(dw, t)
    return dw
