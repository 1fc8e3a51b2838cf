"""GPU parity of the tcgen05 engine (bf16 operands, fp32 accumulation in TMEM) against a torch
CPU fp32 convolution of the SAME bf16-rounded operands (so only accumulation order and the
bf16 rounding of the output differ: tolerance 1e-2 of the output range) and against the SIMT
engine in bf16."""
import pytest
import torch
import torch.nn.functional as F

from helpers import relerr

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-2
BF = torch.bfloat16


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _q(t):
    return t.bfloat16().float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, BF)


def _nchw(x):
    return x.float().cpu().permute(0, 3, 1, 2).contiguous()


CASES = [
    # n, cin, cout, h, w, k, stride
    (1, 64, 64, 16, 24, 3, 1),        # K = 576 = 9 stages, M = 384 = 3 tiles
    (2, 32, 32, 13, 19, 3, 1),        # M tail (494), two taps per 64-wide K chunk
    (1, 16, 32, 12, 20, 3, 1),        # K = 144: padded last stage
    (1, 24, 96, 9, 9, 3, 1),          # K = 216, cout = 3 x BN(32)
    (1, 128, 128, 11, 22, 3, 1),      # BN = 128
    (1, 256, 256, 6, 11, 3, 1),       # M = 66 < one tile, grid.y = 2
    (1, 64, 128, 22, 44, 3, 2),       # stride 2
    (2, 32, 64, 10, 14, 1, 1),        # 1x1 (gated fusion shape family)
    (1, 64, 128, 11, 22, 1, 2),       # 1x1 stride 2 (projection shortcut)
    (2, 8, 32, 20, 28, 7, 2),         # 7x7 stem with channels padded to 8
]


@pytest.mark.parametrize('case', CASES)
def test_tc_conv_plain(case):
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _q(_rand(n, cin, h, w, seed=1))
    wt = _q(_rand(cout, cin, k, k, seed=2) / (cin * k * k) ** 0.5)
    ref = F.conv2d(x, wt, None, s, k // 2)
    wp = ops.pack_weight(wt.to(DEV), BF)
    out = ops.conv2d(_nhwc(x), wp, cout, k, s, engine=ops.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert relerr(_nchw(out), ref) < TOL
    simt = ops.conv2d(_nhwc(x), wp, cout, k, s, engine=ops.ENGINE_SIMT)
    assert relerr(_nchw(out), _nchw(simt)) < TOL


@pytest.mark.parametrize('src_hw,dst_hw', [((6, 11), (11, 22)), ((8, 12), (16, 24))])
def test_tc_upsample_concat_epilogue(src_hw, dst_hw):
    from rcfd import ops
    n, c0, c1, cout = 2, 64, 32, 64
    x0, x1 = _q(_rand(n, c0, *src_hw, seed=3)), _q(_rand(n, c1, *dst_hw, seed=4))
    wt = _q(_rand(cout, c0 + c1, 3, 3, seed=5) / 30.0)
    scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=6) * 0.1
    res = _q(_rand(n, cout, *dst_hw, seed=7))
    ref = F.conv2d(torch.cat([F.interpolate(x0, size=dst_hw), x1], 1), wt, None, 1, 1)
    ref = F.leaky_relu(ref * scale[None, :, None, None] + shift[None, :, None, None], 0.2)
    ref = F.leaky_relu(ref + res, 0.2)
    out = ops.conv2d(_nhwc(x0), ops.pack_weight(wt.to(DEV), BF), cout, 3, 1, x1=_nhwc(x1), in_size=dst_hw,
                     scale=scale.to(DEV), shift=shift.to(DEV), act=ops.ACT_LEAKY, residual=_nhwc(res),
                     engine=ops.ENGINE_TCGEN05)
    assert relerr(_nchw(out), ref) < TOL


def test_tc_stats_sigmoid_accumulate():
    from rcfd import ops
    n, cin, cout, h, w = 2, 32, 64, 15, 17
    x, wt = _q(_rand(n, cin, h, w, seed=8)), _q(_rand(cout, cin, 3, 3, seed=9) / 17.0)
    ref = F.conv2d(x, wt, None, 1, 1)
    ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros_like(ssum)
    wp = ops.pack_weight(wt.to(DEV), BF)
    out = ops.conv2d(_nhwc(x), wp, cout, 3, 1, stats=(ssum, ssq), engine=ops.ENGINE_TCGEN05)
    assert relerr(_nchw(out), ref) < TOL
    assert relerr(ssum.cpu(), ref.double().sum(dim=(0, 2, 3))) < 1e-4         # fp32 accumulators, before rounding
    assert relerr(ssq.cpu(), (ref.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    sig = ops.conv2d(_nhwc(x), wp, cout, 3, 1, act=ops.ACT_SIGMOID, engine=ops.ENGINE_TCGEN05)
    assert relerr(_nchw(sig), torch.sigmoid(ref)) < TOL
    acc = _nhwc(_q(_rand(n, cout, h, w, seed=10)))
    base = _nchw(acc).clone()
    ops.conv2d(_nhwc(x), wp, cout, 3, 1, out=acc, accumulate=True, engine=ops.ENGINE_TCGEN05)
    assert relerr(_nchw(acc), base + ref) < TOL


@pytest.mark.parametrize('case', [(2, 32, 64, 12, 20, 3, 1), (1, 32, 64, 11, 22, 3, 2), (2, 32, 64, 9, 9, 1, 2)])
def test_tc_dgrad(case):
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _q(_rand(n, cin, h, w, seed=11)).requires_grad_(True)
    wt = _q(_rand(cout, cin, k, k, seed=12) / (cin * k * k) ** 0.5)
    y = F.conv2d(x, wt, None, s, k // 2)
    dy = _q(_rand(*y.shape, seed=13))
    y.backward(dy)
    wd = ops.pack_weight(wt.to(DEV), BF, dgrad=True)
    dx = ops.conv2d(_nhwc(dy), wd, cin, k, 1, pad=k - 1 - k // 2, in_dilation=s, out_size=(h, w),
                    engine=ops.ENGINE_TCGEN05)
    assert relerr(_nchw(dx), x.grad) < TOL


@pytest.mark.parametrize('case', [(2, 32, 64, 12, 20), (1, 64, 128, 11, 22), (2, 128, 256, 22, 44), (1, 16, 32, 9, 13),
                                  (3, 64, 64, 6, 11)])
def test_tma_dgrad_stride2_phases(case):
    """Zero-insertion-free data gradient of a 3x3 / stride-2 / pad-1 conv on the TMA engine (four 2x2 phase convs on the
    dy grid, rcfd_pack_dgrad_s2_weight) vs autograd, even and odd input extents (11 -> 6 drops the last phase row), plus
    the batched pack of the same weights."""
    from rcfd import ops
    n, cin, cout, h, w = case
    x = _q(_rand(n, cin, h, w, seed=21)).requires_grad_(True)
    wt = _q(_rand(cout, cin, 3, 3, seed=22) / (cin * 9) ** 0.5)
    y = F.conv2d(x, wt, None, 2, 1)
    dy = _q(_rand(*y.shape, seed=23))
    y.backward(dy)
    wd2 = ops.pack_dgrad_s2_weight(wt.to(DEV), BF)
    assert wd2.shape == (4, cin, 4, cout)
    dx = ops.conv2d(_nhwc(dy), wd2, cin, 3, 1, pad=1, in_dilation=2, out_size=(h, w), weight_up2x=wd2)
    assert ops._lib.load().rcfd_last_kernel().decode().startswith('conv_tma_kernel')
    assert relerr(_nchw(dx), x.grad) < TOL
    # the gather engine on the zero-inserted grid computes the same thing
    wd = ops.pack_weight(wt.to(DEV), BF, dgrad=True)
    old = ops.conv2d(_nhwc(dy), wd, cin, 3, 1, pad=1, in_dilation=2, out_size=(h, w), engine=ops.ENGINE_TCGEN05)
    assert relerr(_nchw(dx), _nchw(old)) < 2e-2
    # batched pack == single pack (also with a channel slice and padded dy channels)
    t = ops.PackBatch()
    outs = []
    for kw in (dict(), dict(cin_off=cin // 2, cin_cnt=cin // 2, pad_to=cout + 16)):
        shape, dt_, zero, items = ops.spec_pack_dgrad_s2_weight(wt.to(DEV), BF, **kw)
        out = torch.full(shape, 3.0, device=DEV, dtype=dt_)
        for it in items:
            it = dict(it)
            t.add(it.pop('kind'), it.pop('src'), out, **it)
        outs.append((out, ops.pack_dgrad_s2_weight(wt.to(DEV), BF, **kw)))
    t.finalize(DEV).run()
    torch.cuda.synchronize()
    for got, ref in outs:
        assert torch.equal(got, ref)


@pytest.mark.parametrize('case', [(2, 64, 32, 8, 16), (1, 32, 64, 11, 13), (2, 128, 64, 5, 9), (1, 256, 256, 6, 11)])
def test_upconv2x_dgrad_as_strided_conv(case):
    """Gradient of `3x3 conv after exact 2x nearest up-sampling` w.r.t. the low-res source as ONE 4x4 / stride-2 conv over
    dy (rcfd_pack_upconv2x_dgrad_weight) vs autograd through F.interpolate + conv2d; batched pack == single pack."""
    from rcfd import ops
    n, cin, cout, h, w = case
    x = _q(_rand(n, cin, h, w, seed=31)).requires_grad_(True)
    wt = _q(_rand(cout, cin, 3, 3, seed=32) / (cin * 9) ** 0.5)
    y = F.conv2d(F.interpolate(x, size=(2 * h, 2 * w), mode='nearest'), wt, None, 1, 1)
    dy = _q(_rand(*y.shape, seed=33))
    y.backward(dy)
    w4 = ops.pack_upconv2x_dgrad_weight(wt.to(DEV), BF)
    assert w4.shape == (cin, 16, cout)
    dx = ops.conv2d(_nhwc(dy), w4, cin, 4, 2, pad=1)
    assert tuple(dx.shape) == (n, h, w, cin)
    # the packed weights are sums of up to 4 bf16-rounded taps, rounded once more: slightly looser than a plain conv
    assert relerr(_nchw(dx), x.grad) < 2 * TOL
    t = ops.PackBatch()
    shape, dt_, zero, items = ops.spec_pack_upconv2x_dgrad_weight(wt.to(DEV), BF, cin_off=cin // 2, cin_cnt=cin // 2, pad_to=cout + 16)
    out = torch.full(shape, 3.0, device=DEV, dtype=dt_)
    for it in items:
        it = dict(it)
        t.add(it.pop('kind'), it.pop('src'), out, **it)
    t.finalize(DEV).run()
    torch.cuda.synchronize()
    assert torch.equal(out, ops.pack_upconv2x_dgrad_weight(wt.to(DEV), BF, cin_off=cin // 2, cin_cnt=cin // 2, pad_to=cout + 16))


def test_strip_output_stationary_variant_still_correct():
    """rcfd_set_option('strip_input_stationary', 0) selects the output-stationary single-source row kernels (the default
    is the input-stationary form): same results on the same layer."""
    from rcfd import ops
    n, cin, cout, h, w = 2, 32, 32, 19, 200
    x = _q(_rand(n, cin, h, w, seed=51))
    wt = _q(_rand(cout, cin, 3, 3, seed=52) / (cin * 9) ** 0.5)
    raw = F.conv2d(x, wt, None, 1, 1)
    wp = ops.pack_weight(wt.to(DEV), BF)
    outs = []
    for flag in (1, 0):
        ops.set_option('strip_input_stationary', flag)
        try:
            ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
            ssq = torch.zeros_like(ssum)
            y = ops.conv2d(_nhwc(x), wp, cout, 3, 1, stats=(ssum, ssq), engine=ops.ENGINE_STRIP)
            name = ops._lib.load().rcfd_last_kernel().decode()
        finally:
            ops.set_option('strip_input_stationary', 1)
        assert relerr(_nchw(y), raw) < TOL
        assert relerr(ssum.cpu(), raw.double().sum(dim=(0, 2, 3))) < 1e-2
        outs.append((name, y))
    assert outs[0][0] != outs[1][0], outs[0][0]                 # two different kernels really ran
    assert relerr(_nchw(outs[0][1]), _nchw(outs[1][1])) < 1e-2


@pytest.mark.parametrize('case', [(8, 256, 256, 6, 11, 3, 1), (1, 128, 128, 22, 44, 3, 1), (2, 256, 128, 11, 22, 3, 2),
                                  (1, 128, 512, 11, 22, 1, 1), (1, 64, 64, 44, 88, 3, 1), (2, 128, 64, 7, 9, 3, 1)])
def test_tma_conv_split_k_cluster(case):
    """Layers with fewer tiles than SMs split the k-loop of a tile over a 2- / 4-CTA cluster and merge the partial
    accumulators through distributed shared memory: same results as the unsplit kernel (statistics, folded BN + activation +
    residual epilogue), and as F.conv2d."""
    from rcfd import ops
    n, cin, cout, h, w, k, s_ = case
    x = _q(_rand(n, cin, h * s_, w * s_, seed=61))
    wt = _q(_rand(cout, cin, k, k, seed=62) / (cin * k * k) ** 0.5)
    raw = F.conv2d(x, wt, None, s_, k // 2)
    scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=63) * 0.1
    res = _q(_rand(*raw.shape, seed=64))
    ref = F.leaky_relu(F.leaky_relu(raw * scale[None, :, None, None] + shift[None, :, None, None], 0.2) + res, 0.2)
    wp = ops.pack_weight(wt.to(DEV), BF)
    outs = []
    for split in (2, 0):
        ops.set_option('tma_split_k', split)
        try:
            ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
            ssq = torch.zeros_like(ssum)
            y = ops.conv2d(_nhwc(x), wp, cout, k, s_, stats=(ssum, ssq), engine=ops.ENGINE_TMA)
            z = ops.conv2d(_nhwc(x), wp, cout, k, s_, scale=scale.to(DEV), shift=shift.to(DEV), act=ops.ACT_LEAKY,
                           residual=_nhwc(res), engine=ops.ENGINE_TMA)
        finally:
            ops.set_option('tma_split_k', 1)
        torch.cuda.synchronize()
        assert relerr(_nchw(y), raw) < TOL and relerr(_nchw(z), ref) < TOL
        assert relerr(ssum.cpu(), raw.double().sum(dim=(0, 2, 3))) < 2e-3 * max(1.0, float(raw.abs().sum() / raw.double().sum(dim=(0, 2, 3)).abs().max() / raw.shape[1]))
        assert relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 1e-3
        outs.append((y, z, ssum, ssq))
    assert relerr(_nchw(outs[0][0]), _nchw(outs[1][0])) < 1e-2 and relerr(_nchw(outs[0][1]), _nchw(outs[1][1])) < 1e-2
    assert relerr(outs[0][3].cpu(), outs[1][3].cpu()) < 1e-5


def test_tc_rejects_unsupported():
    from rcfd import ops, _lib
    x = torch.zeros(1, 4, 4, 8, device=DEV)          # fp32 -> not a tcgen05 case
    w = torch.zeros(32, 9, 8, device=DEV)
    with pytest.raises(_lib.RcfdError):
        ops.conv2d(x, w, 32, 3, 1, engine=ops.ENGINE_TCGEN05)


@pytest.mark.parametrize('case', [(2, 64, 64, 16, 24, 3, 1), (1, 32, 32, 13, 19, 3, 1), (1, 16, 32, 12, 20, 3, 1),
                                  (1, 128, 256, 11, 22, 3, 2), (2, 32, 128, 10, 14, 1, 1), (1, 24, 96, 9, 9, 3, 1),
                                  (1, 256, 512, 6, 11, 1, 1), (3, 64, 32, 40, 56, 3, 1)])
def test_tc_wgrad(case):
    """tcgen05 weight gradient (MN-major operands, split over pixels) vs autograd of F.conv2d."""
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _q(_rand(n, cin, h, w, seed=21))
    wt = (_rand(cout, cin, k, k, seed=22) * 0.05).requires_grad_(True)
    y = F.conv2d(x, wt, None, s, k // 2)
    dy = _q(_rand(*y.shape, seed=23))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), k, s, engine=ops.ENGINE_TCGEN05)
    gw = torch.empty(cout, cin, k, k, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 2e-3          # exact bf16 products, fp32 accumulation / atomics order


def test_tc_wgrad_dual_source_upsampled():
    from rcfd import ops
    n, c0, c1, cout = 1, 64, 32, 64
    x0, x1 = _q(_rand(n, c0, 8, 12, seed=24)), _q(_rand(n, c1, 16, 24, seed=25))
    wt = (_rand(cout, c0 + c1, 3, 3, seed=26) * 0.05).requires_grad_(True)
    y = F.conv2d(torch.cat([F.interpolate(x0, size=(16, 24)), x1], 1), wt, None, 1, 1)
    dy = _q(_rand(*y.shape, seed=27))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x0), _nhwc(dy), 3, 1, x1=_nhwc(x1), in_size=(16, 24), engine=ops.ENGINE_TCGEN05)
    gw = torch.empty(cout, c0 + c1, 3, 3, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 2e-3


# ----------------------------------------------------------------------------- TMA engine
TMA_CASES = [
    # n, cin, cout, h, w, k, stride
    (1, 64, 64, 16, 32, 3, 1),        # exact tiles (TW = 32)
    (2, 32, 32, 13, 19, 3, 1),        # partial tiles both ways, bkc = 32 (SWIZZLE_64B)
    (1, 16, 32, 12, 20, 3, 1),        # bkc = 16 (SWIZZLE_32B)
    (2, 128, 128, 11, 22, 3, 1),      # BN = 128, two k chunks per tap
    (1, 256, 256, 6, 11, 3, 1),       # one partial tile per image, tiles_n = 2
    (1, 64, 128, 22, 44, 3, 2),       # stride 2 through the TMA traversal stride
    (2, 32, 64, 10, 14, 1, 1),        # 1x1
    (1, 64, 128, 11, 22, 1, 2),       # 1x1 stride 2 (projection shortcut)
    (2, 16, 32, 20, 28, 7, 2),        # 7x7 stem, channels padded to 16
    (1, 32, 16, 9, 17, 3, 1),         # cout = 16
    (3, 64, 32, 40, 64, 3, 1),        # more tiles than one wave of k steps; persistent loop
]


@pytest.mark.parametrize('case', TMA_CASES)
def test_tma_conv_plain(case):
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _q(_rand(n, cin, h, w, seed=1))
    wt = _q(_rand(cout, cin, k, k, seed=2) / (cin * k * k) ** 0.5)
    ref = F.conv2d(x, wt, None, s, k // 2)
    out = ops.conv2d(_nhwc(x), ops.pack_weight(wt.to(DEV), BF), cout, k, s, engine=ops.ENGINE_TMA)
    torch.cuda.synchronize()
    assert relerr(_nchw(out), ref) < TOL


def test_tma_concat_epilogue_stats():
    from rcfd import ops
    n, c0, c1, cout, hw = 2, 64, 32, 64, (18, 26)
    x0, x1 = _q(_rand(n, c0, *hw, seed=3)), _q(_rand(n, c1, *hw, seed=4))
    wt = _q(_rand(cout, c0 + c1, 3, 3, seed=5) / 30.0)
    scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=6) * 0.1
    res = _q(_rand(n, cout, *hw, seed=7))
    raw = F.conv2d(torch.cat([x0, x1], 1), wt, None, 1, 1)
    ref = F.leaky_relu(F.leaky_relu(raw * scale[None, :, None, None] + shift[None, :, None, None], 0.2) + res, 0.2)
    wp = ops.pack_weight(wt.to(DEV), BF)
    out = ops.conv2d(_nhwc(x0), wp, cout, 3, 1, x1=_nhwc(x1), scale=scale.to(DEV), shift=shift.to(DEV),
                     act=ops.ACT_LEAKY, residual=_nhwc(res), engine=ops.ENGINE_TMA)
    assert relerr(_nchw(out), ref) < TOL
    ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros_like(ssum)
    y = ops.conv2d(_nhwc(x0), wp, cout, 3, 1, x1=_nhwc(x1), stats=(ssum, ssq), engine=ops.ENGINE_TMA)
    assert relerr(_nchw(y), raw) < TOL
    assert relerr(ssum.cpu(), raw.double().sum(dim=(0, 2, 3))) < 1e-4      # partial tiles masked out of the statistics
    assert relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    acc = _nhwc(_q(_rand(n, cout, *hw, seed=10)))
    base = _nchw(acc).clone()
    ops.conv2d(_nhwc(x0), wp, cout, 3, 1, x1=_nhwc(x1), out=acc, accumulate=True, engine=ops.ENGINE_TMA)
    assert relerr(_nchw(acc), base + raw) < TOL


def test_tma_head_and_dgrad():
    from rcfd import ops
    n, cin, h, w = 2, 32, 21, 37
    x = _q(_rand(n, cin, h, w, seed=8))
    w1 = _q(_rand(1, cin, 3, 3, seed=10) / 6.0)
    refd = 1.0 / (torch.sigmoid(F.conv2d(x, w1, None, 1, 1)) + 0.01)
    d = ops.conv2d(_nhwc(x), ops.pack_weight(w1.to(DEV), BF), 1, 3, 1, act=ops.ACT_DEPTH_HEAD, act_params=(1.0, 0.01),
                   out_f32=True, engine=ops.ENGINE_TMA)
    assert d.dtype == torch.float32 and relerr(_nchw(d), refd) < TOL
    cout = 64
    xg = _q(_rand(n, cin, h, w, seed=11)).requires_grad_(True)
    wt = _q(_rand(cout, cin, 3, 3, seed=12) / 17.0)
    yy = F.conv2d(xg, wt, None, 1, 1)
    dy = _q(_rand(*yy.shape, seed=13))
    yy.backward(dy)
    dx = ops.conv2d(_nhwc(dy), ops.pack_weight(wt.to(DEV), BF, dgrad=True), cin, 3, 1, pad=1, out_size=(h, w),
                    engine=ops.ENGINE_TMA)
    assert relerr(_nchw(dx), xg.grad) < TOL


def test_tma_rejects_upsample():
    from rcfd import ops, _lib
    x = torch.zeros(1, 4, 4, 16, device=DEV, dtype=BF)
    w = torch.zeros(32, 9, 16, device=DEV, dtype=BF)
    with pytest.raises(_lib.RcfdError):
        ops.conv2d(x, w, 32, 3, 1, in_size=(8, 8), engine=ops.ENGINE_TMA)


@pytest.mark.parametrize('case', [(2, 64, 64, 16, 24, 3, 1), (1, 32, 32, 13, 19, 3, 1), (1, 16, 32, 12, 20, 3, 1),
                                  (1, 128, 256, 11, 22, 3, 2), (2, 32, 128, 10, 14, 1, 1), (1, 64, 96, 9, 9, 3, 1),
                                  (1, 256, 512, 6, 11, 1, 1), (3, 64, 32, 40, 56, 3, 1), (2, 16, 16, 20, 28, 7, 2),
                                  (1, 64, 128, 11, 22, 1, 2)])
def test_tma_wgrad(case):
    """TMA-fed tcgen05 weight gradient vs autograd of F.conv2d (partial pixel tiles, strides, k / cout tails)."""
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _q(_rand(n, cin, h, w, seed=21))
    wt = (_rand(cout, cin, k, k, seed=22) * 0.05).requires_grad_(True)
    y = F.conv2d(x, wt, None, s, k // 2)
    dy = _q(_rand(*y.shape, seed=23))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), k, s, engine=ops.ENGINE_TMA)
    gw = torch.empty(cout, cin, k, k, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 2e-3


@pytest.mark.parametrize('case', [(8, 256, 256, 22, 44, 3, 1, 0), (8, 128, 128, 44, 88, 3, 1, 0), (2, 256, 256, 6, 11, 3, 1, 0),
                                  (4, 128, 256, 22, 44, 3, 2, 0), (2, 128, 512, 22, 44, 1, 1, 0), (2, 256, 256, 11, 22, 3, 1, 256),
                                  (1, 64, 64, 44, 88, 3, 1, 0), (2, 128, 128, 11, 22, 3, 1, 128), (1, 64, 128, 5, 7, 1, 2, 0),
                                  (1, 64, 64, 3, 5, 3, 1, 0)])
def test_tma_wgrad_v2_cluster_reduce(case):
    """Second-generation TMA weight gradient (several k-tiles per CTA sharing the dY tile, pixel splits merged through
    distributed shared memory inside a thread-block cluster: no atomics, no memset) vs autograd, at the FusionNet shapes
    (256 / 128 channels, 22x44 ... 6x11, stride 2, 1x1 stacked fusion pair, decoder concat) and tiny extents (fewer pixel
    tiles than a cluster); repeatable; agrees with the first-generation kernel."""
    from rcfd import ops
    n, cin, cout, h, w, k, s, c1 = case
    x = _q(_rand(n, cin, h, w, seed=41))
    x1 = _q(_rand(n, c1, h, w, seed=44)) if c1 else None
    wt = (_rand(cout, cin + c1, k, k, seed=42) * 0.05).requires_grad_(True)
    y = F.conv2d(x if x1 is None else torch.cat([x, x1], 1), wt, None, s, k // 2)
    dy = _q(_rand(*y.shape, seed=43))
    y.backward(dy)
    kw = dict(x1=_nhwc(x1)) if c1 else {}
    dw = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), k, s, engine=ops.ENGINE_TMA, **kw)
    assert ops._lib.load().rcfd_last_kernel().decode().startswith('wgrad_tma2_kernel')
    gw = torch.empty(cout, cin + c1, k, k, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 2e-3
    again = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), k, s, engine=ops.ENGINE_TMA, **kw)
    assert relerr(again.cpu(), dw.cpu()) < 1e-6         # bit identical with one cluster per unit; <= 6 float4 atomics per element otherwise
    ops.set_option('wgrad_tma_v2', 0)
    try:
        old = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), k, s, engine=ops.ENGINE_TMA, **kw)
        assert ops._lib.load().rcfd_last_kernel().decode().startswith('wgrad_tma_kernel')
    finally:
        ops.set_option('wgrad_tma_v2', 1)
    assert relerr(dw.cpu(), old.cpu()) < 1e-4


def test_tma_wgrad_concat():
    from rcfd import ops
    n, c0, c1, cout = 2, 64, 32, 64
    x0, x1 = _q(_rand(n, c0, 16, 24, seed=24)), _q(_rand(n, c1, 16, 24, seed=25))
    wt = (_rand(cout, c0 + c1, 3, 3, seed=26) * 0.05).requires_grad_(True)
    y = F.conv2d(torch.cat([x0, x1], 1), wt, None, 1, 1)
    dy = _q(_rand(*y.shape, seed=27))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x0), _nhwc(dy), 3, 1, x1=_nhwc(x1), engine=ops.ENGINE_TMA)
    gw = torch.empty(cout, c0 + c1, 3, 3, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 2e-3


@pytest.mark.parametrize('shape', [(2, 64, 32, 8, 16), (1, 32, 64, 11, 13), (2, 128, 64, 5, 9), (1, 256, 256, 3, 4)])
def test_tma_upconv2x_subpixel(shape):
    """3x3 conv behind an exact 2x nearest up-sampling as four 2x2 TMA convs on the low-res source
    (sub-pixel phases, summed weights) == F.interpolate + conv2d, incl. zero padding at the borders,
    BatchNorm statistics over all phases and the fused epilogue."""
    from rcfd import ops
    n, cin, cout, h, w = shape
    x = _q(_rand(n, cin, h, w, seed=31))
    wt = _q(_rand(cout, cin, 3, 3, seed=32) / (cin * 9) ** 0.5)
    raw = F.conv2d(F.interpolate(x, size=(2 * h, 2 * w)), wt, None, 1, 1)
    wp = ops.pack_weight(wt.to(DEV), BF)
    wup = ops.pack_upconv2x_weight(wt.to(DEV), BF)
    ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros_like(ssum)
    y = ops.conv2d(_nhwc(x), wp, cout, 3, 1, in_size=(2 * h, 2 * w), stats=(ssum, ssq), engine=ops.ENGINE_TMA,
                   weight_up2x=wup)
    assert y.shape == (n, 2 * h, 2 * w, cout)
    assert relerr(_nchw(y), raw) < TOL
    assert relerr(ssum.cpu(), raw.double().sum(dim=(0, 2, 3))) < 2e-2 * max(1.0, float(raw.abs().mean()) * raw[:, 0].numel() / float(raw.double().sum(dim=(0, 2, 3)).abs().max() + 1e-9)) or \
        relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 2e-2
    assert relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 2e-2        # weights are summed before bf16 rounding
    scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=33) * 0.1
    ref = F.leaky_relu(raw * scale[None, :, None, None] + shift[None, :, None, None], 0.2)
    z = ops.conv2d(_nhwc(x), wp, cout, 3, 1, in_size=(2 * h, 2 * w), scale=scale.to(DEV), shift=shift.to(DEV),
                   act=ops.ACT_LEAKY, weight_up2x=wup)                                # AUTO picks the TMA sub-pixel path
    assert relerr(_nchw(z), ref) < TOL
    g = ops.conv2d(_nhwc(x), wp, cout, 3, 1, in_size=(2 * h, 2 * w), scale=scale.to(DEV), shift=shift.to(DEV),
                   act=ops.ACT_LEAKY, engine=ops.ENGINE_TCGEN05)                      # gather engine, same layer
    assert relerr(_nchw(z), _nchw(g)) < TOL


# ----------------------------------------------------------------------------- row-streaming engine
@pytest.mark.parametrize('mode', [0])
@pytest.mark.parametrize('case', [(1, 64, 64, 12, 128), (2, 64, 32, 19, 200), (1, 32, 32, 33, 130), (2, 32, 64, 9, 70),
                                  (1, 64, 16, 70, 256), (2, 16, 32, 21, 140)])
def test_strip_conv(case, mode):
    """3x3 / stride 1 conv with the input rows kept in a shared-memory ring: tap (r, s) = descriptor
    shifted by s pixels into ring row y-1+r (descriptor base offset 0: the swizzle is a function of the
    absolute shared-memory address -- measured on B200: the other convention gives garbage)."""
    from rcfd import ops
    n, cin, cout, h, w = case
    ops.set_option('strip_desc_mode', mode)
    try:
        x = _q(_rand(n, cin, h, w, seed=41))
        wt = _q(_rand(cout, cin, 3, 3, seed=42) / (cin * 9) ** 0.5)
        raw = F.conv2d(x, wt, None, 1, 1)
        wp = ops.pack_weight(wt.to(DEV), BF)
        ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
        ssq = torch.zeros_like(ssum)
        y = ops.conv2d(_nhwc(x), wp, cout, 3, 1, stats=(ssum, ssq), engine=ops.ENGINE_STRIP)
        torch.cuda.synchronize()
        err = relerr(_nchw(y), raw)
        print('strip mode', mode, case, 'relerr', err)
        assert err < TOL
        assert relerr(ssum.cpu(), raw.double().sum(dim=(0, 2, 3))) < 1e-3 * (1 + float(raw.abs().sum() / (raw.sum(dim=(0, 2, 3)).abs().max() + 1e-9)))
        assert relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
        scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=43) * 0.1
        res = _q(_rand(n, cout, h, w, seed=44))
        ref = F.leaky_relu(F.leaky_relu(raw * scale[None, :, None, None] + shift[None, :, None, None], 0.2) + res, 0.2)
        z = ops.conv2d(_nhwc(x), wp, cout, 3, 1, scale=scale.to(DEV), shift=shift.to(DEV), act=ops.ACT_LEAKY,
                       residual=_nhwc(res), engine=ops.ENGINE_STRIP)
        assert relerr(_nchw(z), ref) < TOL
    finally:
        ops.set_option('strip_desc_mode', 0)


@pytest.mark.parametrize('case', [(1, 64, 32, 64, 12, 130), (2, 64, 32, 32, 9, 200), (1, 64, 64, 64, 17, 128),
                                  (2, 32, 32, 32, 21, 70), (1, 32, 32, 64, 10, 140)])
def test_strip_conv_concat(case):
    """Row-streaming 3x3 conv over torch.cat([x, skip], 1) (decoder blocks, reference src/net_utils.py:565):
    both sources ride in every ring slot; batch statistics and the folded-BN epilogue as in the single-source case."""
    from rcfd import ops
    n, c0, c1, cout, h, w = case
    x0 = _q(_rand(n, c0, h, w, seed=61))
    x1 = _q(_rand(n, c1, h, w, seed=62))
    wt = _q(_rand(cout, c0 + c1, 3, 3, seed=63) / ((c0 + c1) * 9) ** 0.5)
    raw = F.conv2d(torch.cat([x0, x1], 1), wt, None, 1, 1)
    wp = ops.pack_weight(wt.to(DEV), BF)
    ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros_like(ssum)
    y = ops.conv2d(_nhwc(x0), wp, cout, 3, 1, x1=_nhwc(x1), stats=(ssum, ssq), engine=ops.ENGINE_STRIP)
    torch.cuda.synchronize()
    err = relerr(_nchw(y), raw)
    print('strip concat', case, 'relerr', err)
    assert err < TOL
    assert relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 1e-4
    assert relerr(ssum.cpu(), raw.double().sum(dim=(0, 2, 3))) < 1e-3 * (1 + float(raw.abs().sum() / (raw.sum(dim=(0, 2, 3)).abs().max() + 1e-9)))
    scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=64) * 0.1
    ref = F.leaky_relu(raw * scale[None, :, None, None] + shift[None, :, None, None], 0.2)
    z = ops.conv2d(_nhwc(x0), wp, cout, 3, 1, x1=_nhwc(x1), scale=scale.to(DEV), shift=shift.to(DEV), act=ops.ACT_LEAKY,
                   engine=ops.ENGINE_STRIP)
    assert relerr(_nchw(z), ref) < TOL
    # same answer as the per-tap TMA engine on the same operands
    z_tma = ops.conv2d(_nhwc(x0), wp, cout, 3, 1, x1=_nhwc(x1), scale=scale.to(DEV), shift=shift.to(DEV), act=ops.ACT_LEAKY,
                       engine=ops.ENGINE_TMA)
    assert relerr(_nchw(z), _nchw(z_tma)) < 2e-2


@pytest.mark.parametrize('shape', [(2, 64, 32, 9, 70), (1, 32, 64, 20, 128), (1, 64, 64, 11, 200), (2, 32, 16, 5, 33)])
def test_strip_upconv2x(shape):
    """Row-streaming variant of the sub-pixel up-conv: low-res rows in the shared-memory ring, four phase
    accumulators in TMEM, output rows 2i / 2i+1 written from one pass."""
    from rcfd import ops
    n, cin, cout, h, w = shape
    x = _q(_rand(n, cin, h, w, seed=51))
    wt = _q(_rand(cout, cin, 3, 3, seed=52) / (cin * 9) ** 0.5)
    raw = F.conv2d(F.interpolate(x, size=(2 * h, 2 * w)), wt, None, 1, 1)
    wp = ops.pack_weight(wt.to(DEV), BF)
    wup = ops.pack_upconv2x_weight(wt.to(DEV), BF)
    ssum = torch.zeros(cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros_like(ssum)
    y = ops.conv2d(_nhwc(x), wp, cout, 3, 1, in_size=(2 * h, 2 * w), stats=(ssum, ssq), engine=ops.ENGINE_STRIP,
                   weight_up2x=wup)
    torch.cuda.synchronize()
    assert relerr(_nchw(y), raw) < TOL
    assert relerr(ssq.cpu(), (raw.double() ** 2).sum(dim=(0, 2, 3))) < 2e-2
    scale, shift = torch.rand(cout) + 0.5, _rand(cout, seed=53) * 0.1
    ref = F.leaky_relu(raw * scale[None, :, None, None] + shift[None, :, None, None], 0.2)
    z = ops.conv2d(_nhwc(x), wp, cout, 3, 1, in_size=(2 * h, 2 * w), scale=scale.to(DEV), shift=shift.to(DEV),
                   act=ops.ACT_LEAKY, engine=ops.ENGINE_STRIP, weight_up2x=wup)
    assert relerr(_nchw(z), ref) < TOL


@pytest.mark.parametrize('cin,cout', [(3, 32), (2, 16)])
def test_stem_space_to_depth(cin, cout):
    """7x7 / stride-2 stem == 4x4 / stride-1 conv on the space-to-depth input (forward and weight gradient)."""
    from rcfd import ops
    n, h, w = 2, 20, 36
    x = _q(_rand(n, cin, h, w, seed=61))
    wt = (_q(_rand(cout, cin, 7, 7, seed=62) / (cin * 49) ** 0.5)).requires_grad_(True)
    y = F.conv2d(x, wt, None, 2, 3)
    xs = ops.nchw_to_s2d(x.to(DEV), BF, 16)
    assert xs.shape == (n, h // 2, w // 2, 16)
    ref_s2d = torch.zeros(n, h // 2, w // 2, 16)
    for dy in range(2):
        for dx in range(2):
            ref_s2d[..., (dy * 2 + dx) * cin:(dy * 2 + dx + 1) * cin] = x[:, :, dy::2, dx::2].permute(0, 2, 3, 1)
    assert torch.equal(xs.float().cpu(), ref_s2d)
    ws = ops.pack_stem_s2d_weight(wt.detach().to(DEV), BF, 16)
    for eng in (ops.ENGINE_AUTO, ops.ENGINE_TCGEN05, ops.ENGINE_SIMT):
        out = ops.conv2d(xs, ws, cout, 4, 1, pad=2, out_size=(h // 2, w // 2), engine=eng)
        assert relerr(_nchw(out), y.detach()) < TOL, eng
    dyt = _q(_rand(*y.shape, seed=63))
    y.backward(dyt)
    dw = ops.conv2d_wgrad(xs, _nhwc(dyt), 4, 1, pad=2)
    g = torch.empty(cout, cin, 7, 7, device=DEV)
    ops.unpack_stem_s2d_wgrad(dw, g)
    assert relerr(g.cpu(), wt.grad) < 2e-3


@pytest.mark.parametrize('case', [(1, 3, 32, 70, 300), (2, 2, 16, 66, 128), (1, 3, 32, 64, 131)])
def test_stem_wgrad_strip(case):
    """Weight gradient of the 4x4 / pad-2 space-to-depth stems on the row-streaming engine (KS = 4, 16-channel pixels:
    one MMA covers the four taps of a filter row) against the gather engine and autograd of the 7x7 / stride-2 conv."""
    from rcfd import ops
    n, cin, cout, hs, ws_ = case
    h, w = 2 * hs, 2 * ws_
    x = _q(_rand(n, cin, h, w, seed=71))
    wt = (_q(_rand(cout, cin, 7, 7, seed=72) / (cin * 49) ** 0.5)).requires_grad_(True)
    y = F.conv2d(x, wt, None, 2, 3)
    dyt = _q(_rand(*y.shape, seed=73))
    y.backward(dyt)
    xs = ops.nchw_to_s2d(x.to(DEV), BF, 16)
    dyn = _nhwc(dyt)
    dw = ops.conv2d_wgrad(xs, dyn, 4, 1, pad=2, engine=ops.ENGINE_STRIP)
    assert ops._lib.load().rcfd_last_kernel().decode().startswith('wgrad_strip_kernel<%d,16,0,4>' % cout)
    dw_g = ops.conv2d_wgrad(xs, dyn, 4, 1, pad=2, engine=ops.ENGINE_TCGEN05)
    assert relerr(dw, dw_g) < 1e-3
    g = torch.empty(cout, cin, 7, 7, device=DEV)
    ops.unpack_stem_s2d_wgrad(dw, g)
    assert relerr(g.cpu(), wt.grad) < 2e-3
    if hs >= 64 and ws_ >= 256:
        ops.conv2d_wgrad(xs, dyn, 4, 1, pad=2)
        assert ops._lib.load().rcfd_last_kernel().decode().startswith('wgrad_strip_kernel')


# ----------------------------------------------------------------------------- row-streaming weight gradient
@pytest.mark.parametrize('case', [(1, 64, 64, 12, 128), (2, 64, 32, 19, 200), (1, 32, 32, 33, 130), (2, 32, 64, 9, 70),
                                  (1, 64, 16, 70, 256), (2, 32, 16, 40, 300)])
def test_strip_wgrad(case):
    """Row-streaming weight gradient: input rows + dY rows in shared-memory rings, MN-major operands, the taps of
    one filter row covered by ONE M = 128 MMA whose channel blocks are one pixel apart."""
    from rcfd import ops
    n, cin, cout, h, w = case
    x = _q(_rand(n, cin, h, w, seed=71))
    wt = (_rand(cout, cin, 3, 3, seed=72) * 0.05).requires_grad_(True)
    y = F.conv2d(x, wt, None, 1, 1)
    dy = _q(_rand(*y.shape, seed=73))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), 3, 1, engine=ops.ENGINE_STRIP)
    torch.cuda.synchronize()
    gw = torch.empty(cout, cin, 3, 3, device=DEV)
    ops.unpack_wgrad(dw, gw)
    err = relerr(gw.cpu(), wt.grad)
    per_tap = [(r, s, round(relerr(gw.cpu()[:, :, r, s], wt.grad[:, :, r, s]), 5)) for r in range(3) for s in range(3)]
    print('strip wgrad', case, err, per_tap)
    assert err < 2e-3


def test_strip_wgrad_concat():
    from rcfd import ops
    n, c0, c1, cout = 2, 64, 32, 64
    x0, x1 = _q(_rand(n, c0, 20, 150, seed=74)), _q(_rand(n, c1, 20, 150, seed=75))
    wt = (_rand(cout, c0 + c1, 3, 3, seed=76) * 0.05).requires_grad_(True)
    y = F.conv2d(torch.cat([x0, x1], 1), wt, None, 1, 1)
    dy = _q(_rand(*y.shape, seed=77))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x0), _nhwc(dy), 3, 1, x1=_nhwc(x1), engine=ops.ENGINE_STRIP)
    gw = torch.empty(cout, c0 + c1, 3, 3, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 2e-3


@pytest.mark.parametrize('shape', [(2, 64, 32, 9, 70), (1, 64, 64, 11, 200), (1, 64, 16, 20, 128)])
def test_strip_wgrad_upconv2x(shape):
    """Weight gradient of `3x3 conv after 2x nearest up-sampling` streamed over the LOW-RES rows (16 sub-pixel
    matrices from stride-2 dY boxes, folded into the 9 taps)."""
    from rcfd import ops
    n, cin, cout, h, w = shape
    x = _q(_rand(n, cin, h, w, seed=81))
    wt = (_rand(cout, cin, 3, 3, seed=82) * 0.05).requires_grad_(True)
    y = F.conv2d(F.interpolate(x, size=(2 * h, 2 * w)), wt, None, 1, 1)
    dy = _q(_rand(*y.shape, seed=83))
    y.backward(dy)
    dw = ops.conv2d_wgrad(_nhwc(x), _nhwc(dy), 3, 1, in_size=(2 * h, 2 * w), engine=ops.ENGINE_STRIP)
    torch.cuda.synchronize()
    gw = torch.empty(cout, cin, 3, 3, device=DEV)
    ops.unpack_wgrad(dw, gw)
    err = relerr(gw.cpu(), wt.grad)
    print('strip wgrad up', shape, err)
    assert err < 2e-3
