"""GPU parity: the implicit-GEMM convolution (C-ABI rcfd_conv2d_fwd / _wgrad) against
torch CPU fp32 references of the reference's call sites (src/net_utils.py:63-91,196,565)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import relerr

pytestmark = pytest.mark.gpu

F32_TOL = 2e-5       # fp32 FMA vs CPU fp32: summation-order noise only
BF16_TOL = 2e-2      # bf16 storage of inputs/weights/outputs (8-bit mantissa), fp32 accumulate


def _dev():
    return torch.device('cuda:0')


def _nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(_dev(), dtype)


def _nchw(x):
    return x.float().cpu().permute(0, 3, 1, 2).contiguous()


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


CASES = [
    # n, cin, cout, h, w, k, stride
    (2, 16, 32, 12, 20, 3, 1),
    (1, 32, 64, 11, 22, 3, 2),      # odd extent, stride 2 -> 6 x 11
    (2, 64, 32, 9, 7, 1, 1),
    (1, 32, 48, 10, 14, 1, 2),
    (2, 3, 32, 20, 28, 7, 2),       # image stem (scalar gather path)
    (2, 2, 16, 20, 28, 7, 2),       # depth stem
    (1, 32, 1, 16, 24, 3, 1),       # output head, cout = 1
    (1, 256, 256, 6, 11, 3, 1),     # latent-size map
    (1, 8, 72, 5, 5, 3, 1),         # cout not a multiple of the tile
]


@pytest.mark.parametrize('dtype,tol', [(torch.float32, F32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize('case', CASES)
def test_conv_plain(case, dtype, tol):
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _rand(n, cin, h, w, seed=1)
    wt = _rand(cout, cin, k, k, seed=2) / (cin * k * k) ** 0.5
    if dtype == torch.bfloat16:        # compare against the same rounded operands
        x, wt = x.bfloat16().float(), wt.bfloat16().float()
    ref = F.conv2d(x, wt, None, s, k // 2)
    wp = ops.pack_weight(wt.to(_dev()), dtype)
    out = ops.conv2d(_nhwc(x, dtype), wp, cout, k, s, engine=ops.ENGINE_SIMT)
    assert relerr(_nchw(out), ref) < tol


@pytest.mark.parametrize('dtype,tol', [(torch.float32, F32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize('src_hw,dst_hw', [((6, 11), (11, 22)), ((5, 9), (10, 18)), ((4, 4), (7, 9))])
def test_conv_upsample_concat_epilogue(src_hw, dst_hw, dtype, tol):
    """nearest up-sample folded into the loads (incl. the non-integer 6 -> 11 ratio), concat as a
    two-source K loop, folded BN + leaky + residual + leaky in the epilogue."""
    from rcfd import ops
    n, c0, c1, cout = 2, 16, 8, 24
    x0 = _rand(n, c0, *src_hw, seed=3)
    x1 = _rand(n, c1, *dst_hw, seed=4)
    wt = _rand(cout, c0 + c1, 3, 3, seed=5) / 12.0
    scale = torch.rand(cout) + 0.5
    shift = _rand(cout, seed=6) * 0.1
    res = _rand(n, cout, *dst_hw, seed=7)
    if dtype == torch.bfloat16:
        x0, x1, wt, res = [t.bfloat16().float() for t in (x0, x1, wt, res)]
    up = F.interpolate(x0, size=dst_hw)
    ref = F.conv2d(torch.cat([up, x1], 1), wt, None, 1, 1)
    ref = F.leaky_relu(ref * scale[None, :, None, None] + shift[None, :, None, None], 0.2)
    ref = F.leaky_relu(ref + res, 0.2)
    wp = ops.pack_weight(wt.to(_dev()), dtype)
    out = ops.conv2d(_nhwc(x0, dtype), wp, cout, 3, 1, x1=_nhwc(x1, dtype), in_size=dst_hw,
                     scale=scale.to(_dev()), shift=shift.to(_dev()), act=ops.ACT_LEAKY,
                     residual=_nhwc(res, dtype), engine=ops.ENGINE_SIMT)
    assert relerr(_nchw(out), ref) < tol


def test_conv_stats_and_head():
    from rcfd import ops
    n, cin, cout, h, w = 2, 16, 32, 9, 13
    x = _rand(n, cin, h, w, seed=8)
    wt = _rand(cout, cin, 3, 3, seed=9) / 12.0
    ref = F.conv2d(x, wt, None, 1, 1)
    ssum = torch.zeros(cout, dtype=torch.float64, device=_dev())
    ssq = torch.zeros_like(ssum)
    out = ops.conv2d(_nhwc(x, torch.float32), ops.pack_weight(wt.to(_dev()), torch.float32), cout, 3, 1,
                     stats=(ssum, ssq), engine=ops.ENGINE_SIMT)
    assert relerr(_nchw(out), ref) < F32_TOL
    assert relerr(ssum.cpu(), ref.double().sum(dim=(0, 2, 3))) < 1e-5
    assert relerr(ssq.cpu(), (ref.double() ** 2).sum(dim=(0, 2, 3))) < 1e-5
    # depth head epilogue (src/fusionnet_model.py:162-165)
    w1 = _rand(1, cin, 3, 3, seed=10) / 4.0
    logits = F.conv2d(x, w1, None, 1, 1)
    refd = 1.0 / (torch.sigmoid(logits) + 0.01)
    d = ops.conv2d(_nhwc(x, torch.float32), ops.pack_weight(w1.to(_dev()), torch.float32), 1, 3, 1,
                   act=ops.ACT_DEPTH_HEAD, act_params=(1.0, 0.01), out_f32=True, engine=ops.ENGINE_SIMT)
    assert relerr(_nchw(d), refd) < 1e-5


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 5e-5), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize('case', [(2, 16, 32, 12, 20, 3, 1), (1, 32, 64, 11, 22, 3, 2), (2, 16, 24, 9, 9, 1, 2),
                                  (2, 3, 16, 14, 18, 7, 2)])
def test_conv_backward(case, dtype, tol):
    """dgrad (flipped weights; stride 2 through zero insertion) and wgrad vs autograd."""
    from rcfd import ops
    n, cin, cout, h, w, k, s = case
    x = _rand(n, cin, h, w, seed=11)
    wt = _rand(cout, cin, k, k, seed=12) / (cin * k * k) ** 0.5
    if dtype == torch.bfloat16:
        x, wt = x.bfloat16().float(), wt.bfloat16().float()
    x.requires_grad_(True)
    wt.requires_grad_(True)
    y = F.conv2d(x, wt, None, s, k // 2)
    dy = _rand(*y.shape, seed=13)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y.backward(dy)
    xd, dyd = _nhwc(x.detach(), dtype), _nhwc(dy, dtype)
    dw = ops.conv2d_wgrad(xd, dyd, k, s)
    gw = torch.empty(cout, cin, k, k, device=_dev())
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < tol
    if cin % 4 == 0:
        wd = ops.pack_weight(wt.detach().to(_dev()), dtype, dgrad=True)
        dx = ops.conv2d(dyd, wd, cin, k, 1, pad=k - 1 - k // 2, in_dilation=s, out_size=(h, w), engine=ops.ENGINE_SIMT)
        assert relerr(_nchw(dx), x.grad) < tol


def test_wgrad_dual_source_upsampled():
    from rcfd import ops
    n, c0, c1, cout = 1, 8, 12, 16
    x0 = _rand(n, c0, 5, 6, seed=14).requires_grad_(True)
    x1 = _rand(n, c1, 10, 12, seed=15).requires_grad_(True)
    wt = (_rand(cout, c0 + c1, 3, 3, seed=16) / 10).requires_grad_(True)
    y = F.conv2d(torch.cat([F.interpolate(x0, size=(10, 12)), x1], 1), wt, None, 1, 1)
    dy = _rand(*y.shape, seed=17)
    y.backward(dy)
    f32 = torch.float32
    dw = ops.conv2d_wgrad(_nhwc(x0.detach(), f32), _nhwc(dy, f32), 3, 1, x1=_nhwc(x1.detach(), f32), in_size=(10, 12))
    gw = torch.empty(cout, c0 + c1, 3, 3, device=_dev())
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wt.grad) < 5e-5
    # dgrad per source + nearest-upsample backward
    wd0 = ops.pack_weight(wt.detach().to(_dev()), f32, cin_off=0, cin_cnt=c0, dgrad=True)
    dup = ops.conv2d(_nhwc(dy, f32), wd0, c0, 3, 1, engine=ops.ENGINE_SIMT)
    dx0 = ops.upsample_nearest_bwd(dup, (5, 6))
    assert relerr(_nchw(dx0), x0.grad) < 5e-5
    wd1 = ops.pack_weight(wt.detach().to(_dev()), f32, cin_off=c0, cin_cnt=c1, dgrad=True)
    dx1 = ops.conv2d(_nhwc(dy, f32), wd1, c1, 3, 1, engine=ops.ENGINE_SIMT)
    assert relerr(_nchw(dx1), x1.grad) < 5e-5


def test_bad_arguments_raise():
    from rcfd import ops, _lib
    x = torch.zeros(1, 4, 4, 8, device=_dev())
    w = torch.zeros(8, 9, 8, device=_dev())
    with pytest.raises(_lib.RcfdError):
        ops.conv2d(x, w, 8, 3, 1, engine=99)
