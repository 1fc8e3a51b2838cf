"""GPU parity of the HBM-bound passes around the convolutions vs torch CPU fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import relerr

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _nhwc(x, dtype=torch.float32):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)


def _nchw(x):
    return x.float().cpu().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize('act,name', [(1, 'leaky'), (2, 'sigmoid'), (0, 'none')])
def test_batchnorm_train_forward_backward(act, name):
    """finalize + apply + backward vs torch BatchNorm2d(train) -> activation (-> residual, leaky)."""
    from rcfd import ops
    n, c, h, w = 3, 16, 7, 9
    y = (_rand(n, c, h, w, seed=1) * 2 + 0.5).requires_grad_(True)
    res = _rand(n, c, h, w, seed=2).requires_grad_(True) if act == 1 else None
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c) + 0.5)
        bn.bias.copy_(_rand(c, seed=3) * 0.1)
        bn.running_mean.copy_(_rand(c, seed=4) * 0.1)
        bn.running_var.copy_(torch.rand(c) + 0.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    bn.train()
    t = bn(y)
    t = F.leaky_relu(t, 0.2) if act == 1 else (torch.sigmoid(t) if act == 2 else t)
    z = F.leaky_relu(t + res, 0.2) if res is not None else t
    dz = _rand(n, c, h, w, seed=5)
    z.backward(dz)

    yd = _nhwc(y.detach())
    ssum = yd.double().sum(dim=(0, 1, 2))
    ssq = (yd.double() ** 2).sum(dim=(0, 1, 2))
    gamma, beta = bn.weight.detach().to(DEV), bn.bias.detach().to(DEV)
    rm, rv = rm0.to(DEV), rv0.to(DEV)
    scale, shift, mean, invstd = [torch.empty(c, device=DEV) for _ in range(4)]
    ops.bn_finalize(ssum, ssq, gamma, beta, rm, rv, scale, shift, mean, invstd, n * h * w)
    assert relerr(rm.cpu(), bn.running_mean) < 1e-5 and relerr(rv.cpu(), bn.running_var) < 1e-5
    zd = ops.bn_act(yd, scale, shift, act, residual=_nhwc(res.detach()) if res is not None else None)
    assert relerr(_nchw(zd), z.detach()) < 1e-5
    dzd = _nhwc(dz)
    if res is not None:
        dzd = ops.leaky_bwd(dzd, zd)
        assert relerr(_nchw(dzd), res.grad) < 1e-5
    dg, db = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    dy = ops.bn_act_bwd(dzd, yd, scale, shift, mean, invstd, act, dg, db)
    assert relerr(_nchw(dy), y.grad) < 2e-4
    assert relerr(dg.cpu(), bn.weight.grad) < 1e-4 and relerr(db.cpu(), bn.bias.grad) < 1e-4


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('pixels,c', [(528, 256), (1936, 128), (7744, 64), (2049, 32), (37, 8)])
def test_batchnorm_backward_single_launch(dtype, pixels, c, monkeypatch):
    """rcfd_bn_act_bwd_fused (small maps: one CTA per 16-byte channel vector, reduce + apply in one launch, optionally
    behind the post-add LeakyReLU of a ResNetBlock) == leaky_bwd + reduce + apply; the float form also against the
    formulas in float64."""
    from rcfd import ops
    g = torch.Generator().manual_seed(pixels + c)
    y = (torch.randn(1, 1, pixels, c, generator=g) * 1.5 + 0.3).to(DEV, dtype)
    dz = torch.randn(1, 1, pixels, c, generator=g).to(DEV, dtype)
    zpost = torch.randn(1, 1, pixels, c, generator=g).to(DEV, dtype)
    scale = (torch.rand(c, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(c, generator=g) * 0.2).to(DEV)
    mean = y.float().reshape(-1, c).mean(0)
    invstd = 1.0 / torch.sqrt(y.float().reshape(-1, c).var(0, unbiased=False) + 1e-5)
    for act in (ops.ACT_LEAKY, ops.ACT_NONE):
        for post in (None, zpost):
            outs = []
            for limit in (1 << 20, 0):
                monkeypatch.setattr(ops, 'BN_BWD_FUSED_MAX_PIXELS', limit)
                dg, db = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
                r = ops.bn_act_bwd(dz, y, scale, shift, mean, invstd, act, dg, db, post_z=post)
                dy, dzm = r if post is not None else (r, None)
                outs.append((dy, dzm, dg, db))
            (dy1, dzm1, dg1, db1), (dy0, dzm0, dg0, db0) = outs
            if post is not None:
                assert torch.equal(dzm1, dzm0)
            assert relerr(dg1, dg0) < 5e-5 and relerr(db1, db0) < 5e-5
            assert relerr(dy1.float(), dy0.float()) < (1e-2 if dtype == torch.bfloat16 else 1e-5)
            if dtype == torch.float32:
                d = dz.double().reshape(-1, c)
                if post is not None:
                    d = torch.where(post.double().reshape(-1, c) > 0, d, 0.2 * d)
                yd = y.double().reshape(-1, c)
                pre = yd * scale.double() + shift.double()
                if act == ops.ACT_LEAKY:
                    d = torch.where(pre > 0, d, 0.2 * d)
                xhat = (yd - mean.double()) * invstd.double()
                ref = scale.double() * (d - d.mean(0) - xhat * (d * xhat).mean(0))
                assert relerr(dy1.reshape(-1, c), ref.float()) < 1e-5
                assert relerr(dg1, (d * xhat).sum(0).float()) < 1e-5 and relerr(db1, d.sum(0).float()) < 1e-5


def test_bn_fold_and_gate():
    from rcfd import ops
    c, n, h, w = 8, 2, 5, 6
    gamma, beta, rm, rv = torch.rand(c) + 0.5, _rand(c, seed=1), _rand(c, seed=2), torch.rand(c) + 0.5
    scale, shift = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    ops.bn_fold(gamma.to(DEV), beta.to(DEV), rm.to(DEV), rv.to(DEV), scale, shift)
    ref_s = gamma / torch.sqrt(rv + 1e-5)
    assert relerr(scale.cpu(), ref_s) < 1e-6 and relerr(shift.cpu(), beta - rm * ref_s) < 1e-5
    # gate: sigmoid(a) * b + img, forward and backward
    y = _rand(n, 2 * c, h, w, seed=3).requires_grad_(True)
    img = _rand(n, c, h, w, seed=4)
    s2, b2 = torch.rand(2 * c) + 0.5, _rand(2 * c, seed=5) * 0.1
    t = y * s2[None, :, None, None] + b2[None, :, None, None]
    t.retain_grad()
    out = torch.sigmoid(t[:, :c]) * t[:, c:] + img
    dout = _rand(n, c, h, w, seed=6)
    out.backward(dout)
    o = ops.gate_fuse(_nhwc(y.detach()), s2.to(DEV), b2.to(DEV), _nhwc(img))
    assert relerr(_nchw(o), out.detach()) < 1e-5
    dz = ops.gate_fuse_bwd(_nhwc(dout), _nhwc(y.detach()), s2.to(DEV), b2.to(DEV))
    assert relerr(_nchw(dz), t.grad) < 1e-5


@pytest.mark.parametrize('hw', [(8, 12), (11, 7), (5, 5)])
def test_maxpool_forward_backward(hw):
    from rcfd import ops
    x = _rand(2, 8, *hw, seed=7)
    x[0, 0, 0, 0] = x[0, 0, 0, 1] = x[0, 0, 1, 0] = 5.0        # ties -> first maximum
    x.requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = _rand(*y.shape, seed=8)
    y.backward(dy)
    yd = ops.maxpool3x3s2(_nhwc(x.detach()))
    assert torch.equal(_nchw(yd), y.detach())
    dx = ops.maxpool3x3s2_bwd(_nhwc(x.detach()), _nhwc(dy))
    assert relerr(_nchw(dx), x.grad) < 1e-6
    # training variant: recorded arg-max positions instead of re-scanning the input
    yi, idx = ops.maxpool3x3s2_idx(_nhwc(x.detach()))
    assert torch.equal(_nchw(yi), y.detach())
    dxi = ops.maxpool3x3s2_bwd_idx(_nhwc(dy), idx, hw)
    assert torch.equal(dxi, dx)


@pytest.mark.parametrize('src,dst', [((6, 11), (11, 22)), ((5, 9), (10, 18)), ((3, 4), (10, 9))])
def test_upsample_backward(src, dst):
    from rcfd import ops
    x = _rand(2, 4, *src, seed=9).requires_grad_(True)
    up = F.interpolate(x, size=dst)
    d = _rand(*up.shape, seed=10)
    up.backward(d)
    dx = ops.upsample_nearest_bwd(_nhwc(d), src)
    assert relerr(_nchw(dx), x.grad) < 1e-6


@pytest.mark.parametrize('hw', [(8, 12), (11, 7), (6, 9)])
def test_maxpool_upsample_backward_bf16(hw):
    """bf16 16-byte-vector paths: 2x2-block max-pool backward (ties -> first maximum) and exact-2x up-sampling backward."""
    from rcfd import ops
    bf = torch.bfloat16
    x = _rand(2, 16, *hw, seed=12).bfloat16().float()
    x[0, 0, 0, 0] = x[0, 0, 0, 1] = x[0, 0, 1, 0] = 5.0
    x[1, 3, 2, 2] = x[1, 3, 2, 3] = x[1, 3, 3, 2] = x[1, 3, 3, 3] = 7.0
    x.requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = _rand(*y.shape, seed=13).bfloat16().float()
    y.backward(dy)
    xd = x.detach().permute(0, 2, 3, 1).contiguous().to(DEV, bf)
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV, bf)
    dx = ops.maxpool3x3s2_bwd(xd, dyd)
    assert relerr(_nchw(dx), x.grad) < 1e-2          # sums of up to 4 bf16 values rounded to bf16
    assert torch.equal(_nchw(dx) != 0, x.grad != 0)  # the SAME winners
    yi, idx = ops.maxpool3x3s2_idx(xd)
    assert torch.equal(_nchw(yi), y.detach())
    assert torch.equal(ops.maxpool3x3s2_bwd_idx(dyd, idx, hw), dx)
    u = _rand(2, 16, 2 * hw[0], 2 * hw[1], seed=14).bfloat16().float()
    ref = u.view(2, 16, hw[0], 2, hw[1], 2).sum(dim=(3, 5))
    du = ops.upsample_nearest_bwd(u.permute(0, 2, 3, 1).contiguous().to(DEV, bf), hw)
    assert relerr(_nchw(du), ref) < 1e-2


def test_layout_loss_outlier_adam():
    from rcfd import ops
    import fusionnet_oracle as fo
    x = _rand(2, 3, 6, 10, seed=11)
    nh = ops.nchw_to_nhwc(x.to(DEV), torch.float32)
    assert torch.equal(nh.cpu(), x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(nh).cpu(), x)
    # masked L1 (value + gradient) vs the oracle formula
    from rcfd import synth
    out = (torch.rand(2, 1, 16, 24) * 50 + 1).requires_grad_(True)
    gt, lidar = synth.training_targets(2, 16, 24, 5)
    loss_ref = fo.fusionnet_loss(out, gt, lidar, 2.0, 'l1')
    loss_ref.backward()
    loss, dout = ops.masked_l1_loss(out.detach().to(DEV), gt.to(DEV), lidar.to(DEV), 2.0)
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * float(loss_ref)
    assert relerr(dout.cpu(), out.grad) < 1e-5
    # outlier removal
    sparse = gt.clone()
    sparse[0, 0, 3, 3] = 70.0
    sparse[0, 0, 3, 4] = 2.0
    ref = fo.outlier_removal(sparse, 7, 1.5)
    got = ops.outlier_removal(sparse.to(DEV), 7, 1.5)
    assert torch.equal(got.cpu(), ref)
    # depth head backward
    logit = _rand(1, 1, 8, 8, seed=12).requires_grad_(True)
    d = 1.0 / (torch.sigmoid(logit) + 0.01)
    dd = _rand(1, 1, 8, 8, seed=13)
    d.backward(dd)
    dl = ops.depth_head_bwd(dd.to(DEV), d.detach().view(1, 8, 8, 1).to(DEV), 1.0, 0.01, torch.float32, cpad=8)
    assert dl.shape == (1, 8, 8, 8) and float(dl[..., 1:].abs().sum()) == 0.0
    assert relerr(dl[..., 0].cpu().view(1, 1, 8, 8), logit.grad) < 1e-4
    # Adam, 3 steps vs torch.optim.Adam
    p = _rand(1000, seed=14)
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1e-3)
    pd, m, v = p.to(DEV), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    for step in range(1, 4):
        g = _rand(1000, seed=20 + step)
        pt.grad = g.clone()
        opt.step()
        ops.adam_step(pd, g.to(DEV), m, v, 1e-3, 0.9, 0.999, 1e-8, step)
    assert relerr(pd.cpu(), pt.detach()) < 1e-6


def test_padded_pack_unpack_and_fused_adam():
    from rcfd import ops, optim
    w = _rand(16, 3, 7, 7, seed=30).to(DEV)
    pk = ops.pack_weight(w, torch.float32, pad_to=8)
    assert pk.shape == (16, 49, 8) and float(pk[..., 3:].abs().sum()) == 0.0
    assert torch.equal(pk[..., :3].reshape(16, 7, 7, 3).permute(0, 3, 1, 2), w)
    g = torch.zeros_like(w)
    ops.unpack_wgrad(pk, g)                       # padded packed gradient -> OIHW, padding dropped
    assert torch.equal(g, w)
    wd = ops.pack_weight(_rand(1, 32, 3, 3, seed=31).to(DEV), torch.float32, dgrad=True, pad_to=8)
    assert wd.shape == (32, 9, 8) and float(wd[..., 1:].abs().sum()) == 0.0
    # FusedAdam == torch.optim.Adam over several steps, parameters re-pointed into one flat buffer
    ps = [torch.nn.Parameter(_rand(33, 5, seed=40).to(DEV)), torch.nn.Parameter(_rand(7, seed=41).to(DEV))]
    ref = [torch.nn.Parameter(q.detach().clone()) for q in ps]
    fo_, to_ = optim.FusedAdam(ps, lr=1e-3), torch.optim.Adam(ref, lr=1e-3)
    for step in range(3):
        for q, r in zip(ps, ref):
            gq = _rand(*q.shape, seed=50 + step).to(DEV)
            q.grad.copy_(gq)
            r.grad = gq.clone()
        fo_.step()
        to_.step()
    for q, r in zip(ps, ref):
        assert relerr(q.detach().cpu(), r.detach().cpu()) < 1e-6
        assert q.data_ptr() >= fo_.flat_param.data_ptr()


@pytest.mark.parametrize('shape', [(2, 37, 53), (1, 8, 9), (3, 64, 96)])
def test_smoothness_losses_kernels_vs_formulas(shape):
    """fusionnet_losses.smoothness_loss_func / sobel_smoothness_loss_func on CUDA tensors (fused value + gradient kernels)
    == the reference's tensor formulas evaluated on the CPU (the module's own CPU branch, pinned against the reference in
    tests/test_host_logic.py): value, and gradient w.r.t. the prediction through autograd, odd sizes and borders."""
    import fusionnet_losses as L
    n, h, w = shape
    g = torch.Generator().manual_seed(h * w)
    image = torch.rand(n, 3, h, w, generator=g)
    weights = (torch.rand(n, 1, h, w, generator=g) < 0.7).float()
    for kind in ('first_order', (3, 3), (7, 7), (5, 3)):
        p_cpu = (torch.rand(n, 1, h, w, generator=torch.Generator().manual_seed(7)) * 40 + 1).requires_grad_(True)
        p_gpu = p_cpu.detach().clone().to(DEV).requires_grad_(True)
        if kind == 'first_order':
            ref = L.smoothness_loss_func(p_cpu, image)
            got = L.smoothness_loss_func(p_gpu, image.to(DEV))
        else:
            fs = [1, 1, kind[0], kind[1]]
            ref = L.sobel_smoothness_loss_func(p_cpu, image, weights, filter_size=fs)
            got = L.sobel_smoothness_loss_func(p_gpu, image.to(DEV), weights.to(DEV), filter_size=fs)
        (3.0 * ref).backward()
        (3.0 * got).backward()
        assert abs(float(got) - float(ref)) < 1e-5 * abs(float(ref)), kind
        assert relerr(p_gpu.grad.cpu(), p_cpu.grad) < 1e-5, kind
    with torch.no_grad():            # value only (no gradient buffers)
        v = L.smoothness_loss_func(p_gpu.detach(), image.to(DEV))
        assert abs(float(v) - float(L.smoothness_loss_func(p_cpu.detach(), image))) < 1e-5 * abs(float(v))
