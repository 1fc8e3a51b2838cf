"""rcfd_pack_batch: every weight packing / weight-gradient unpacking of a step in one launch == the single-tensor
kernels (bit exact: same index maps, same roundings), and a training step that uses the batched path produces the
gradients of the per-layer path."""
import pytest
import torch

from helpers import relerr, synth_fusionnet_state

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def make_model(cfg, p, precision):
    import fusionnet_model
    m = fusionnet_model.FusionNetModel(device=DEV, **cfg)
    m.encoder.load_state_dict({k[len('encoder.'):]: v for k, v in p.items() if k.startswith('encoder.')})
    m.decoder.load_state_dict({k[len('decoder.'):]: v for k, v in p.items() if k.startswith('decoder.')})
    m.set_precision(precision)
    return m


def _w(cout, cin, k, seed):
    return (torch.randn(cout, cin, k, k, generator=torch.Generator().manual_seed(seed)) * 0.1).to(DEV)


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_pack_batch_equals_single_kernels(dtype):
    from rcfd import ops
    ws = [_w(64, 32, 3, 1), _w(128, 96, 3, 2), _w(32, 3, 3, 3), _w(256, 128, 1, 4), _w(256, 128, 1, 5), _w(32, 64, 3, 6),
          _w(32, 3, 7, 7), _w(16, 2, 7, 8), _w(1, 32, 3, 9)]
    cases = [
        (ops.spec_pack_weight(ws[0], dtype), ops.pack_weight(ws[0], dtype)),
        (ops.spec_pack_weight(ws[1], dtype, cin_off=64, cin_cnt=32, dgrad=True, pad_to=128),
         ops.pack_weight(ws[1], dtype, cin_off=64, cin_cnt=32, dgrad=True, pad_to=128)),
        (ops.spec_pack_weight(ws[1], dtype, cin_off=0, cin_cnt=64, dgrad=True, pad_to=160),       # zero-padded columns
         ops.pack_weight(ws[1], dtype, cin_off=0, cin_cnt=64, dgrad=True, pad_to=160)),
        (ops.spec_pack_weight(ws[2], dtype, pad_to=16), ops.pack_weight(ws[2], dtype, pad_to=16)),
        (ops.spec_pack_stacked_1x1([ws[3], ws[4]], dtype), ops.pack_weight(torch.cat([ws[3], ws[4]], 0), dtype)),
        (ops.spec_pack_stacked_1x1([ws[3], ws[4]], dtype, dgrad=True),
         ops.pack_weight(torch.cat([ws[3], ws[4]], 0), dtype, dgrad=True)),
        (ops.spec_pack_upconv2x_weight(ws[5], dtype), ops.pack_upconv2x_weight(ws[5], dtype)),
        (ops.spec_pack_stem_s2d_weight(ws[6], dtype, 16), ops.pack_stem_s2d_weight(ws[6], dtype, 16)),
        (ops.spec_pack_stem_s2d_weight(ws[7], dtype, 16), ops.pack_stem_s2d_weight(ws[7], dtype, 16)),
        (ops.spec_pack_weight(ws[8], dtype, dgrad=True, pad_to=16), ops.pack_weight(ws[8], dtype, dgrad=True, pad_to=16)),
    ]
    table, outs = ops.PackBatch(), []
    for (shape, dt_, zero, items), ref in cases:
        assert tuple(shape) == tuple(ref.shape) and dt_ == ref.dtype
        out = torch.zeros(shape, device=DEV, dtype=dt_) if zero else torch.full(shape, 7.0, device=DEV, dtype=dt_)
        for it in items:
            it = dict(it)
            table.add(it.pop('kind'), it.pop('src'), out, **it)
        outs.append(out)
    table.finalize(DEV).run()
    torch.cuda.synchronize()
    for i, (out, (_, ref)) in enumerate(zip(outs, cases)):
        assert torch.equal(out, ref), i


def test_unpack_batch_equals_single_kernels():
    from rcfd import ops
    gen = torch.Generator().manual_seed(0)
    dw0 = torch.randn(64, 9, 96, generator=gen).to(DEV)          # cin 96 = concat of 64 + 32
    dw1 = torch.randn(32, 9, 16, generator=gen).to(DEV)          # 3 real channels stored with 16
    dw2 = torch.randn(512, 1, 128, generator=gen).to(DEV)        # stacked 1x1 pair
    dws = torch.randn(32, 16, 16, generator=gen).to(DEV)         # stem
    g0, g1 = torch.empty(64, 96, 3, 3, device=DEV), torch.empty(32, 3, 3, 3, device=DEV)
    g2a, g2b = torch.empty(256, 128, 1, 1, device=DEV), torch.empty(256, 128, 1, 1, device=DEV)
    gs = torch.empty(32, 3, 7, 7, device=DEV)
    refs = [torch.empty_like(t) for t in (g0, g1, g2a, g2b, gs)]
    ops.unpack_wgrad(dw0, refs[0])
    ops.unpack_wgrad(dw1, refs[1])
    ops.unpack_wgrad(dw2[:256], refs[2])
    ops.unpack_wgrad(dw2[256:], refs[3])
    ops.unpack_stem_s2d_wgrad(dws, refs[4])
    t = ops.PackBatch()
    t.add(ops.UNPACK_CONV, dw0, g0, total=g0.numel(), cout=64, cin=96, taps=9, cin_cnt=96, cpad=96)
    t.add(ops.UNPACK_CONV, dw1, g1, total=g1.numel(), cout=32, cin=3, taps=9, cin_cnt=3, cpad=16)
    t.add(ops.UNPACK_CONV, dw2, g2a, total=g2a.numel(), cout=256, cin=128, taps=1, cin_cnt=128, cpad=128)
    t.add(ops.UNPACK_CONV, dw2, g2b, total=g2b.numel(), cout=256, cin=128, taps=1, cin_cnt=128, cpad=128, src_off=256 * 128)
    t.add(ops.UNPACK_STEM_S2D, dws, gs, total=gs.numel(), cout=32, cin=3, taps=16, cpad=16)
    t.finalize(DEV).run()
    torch.cuda.synchronize()
    for got, ref in zip((g0, g1, g2a, g2b, gs), refs):
        assert torch.equal(got, ref)


def test_batched_step_equals_per_layer_step():
    """Step 1 of a model records the pack / unpack plans and runs the per-layer kernels; step 2 on the same data runs
    ONE pack launch and ONE unpack launch.  Same loss, same gradients (up to the order of the wgrad atomics)."""
    from rcfd import optim, synth, _lib
    cfg = synth.CANONICAL_FUSIONNET
    m = make_model(cfg, synth_fusionnet_state(cfg, 5), precision='bf16')
    m.train()
    opt = optim.FusedAdam(m.parameters(), lr=0.0)
    n, h, w = 2, 96, 160
    image, depth = synth.fusionnet_inputs(n, h, w, 5, 'quasi_dense')
    gt, lidar = synth.training_targets(n, h, w, 5)
    image, depth, gt, lidar = [t.to(DEV) for t in (image, depth, gt, lidar)]
    res = []
    for step in range(3):
        opt.flat_grad.zero_()
        l0 = _lib.launch_count
        d = m.forward(image, depth)
        loss, _ = m.compute_loss(image, d, gt, lidar, 'l1', 0.0, -1, None, 2.0)
        loss.backward()
        torch.cuda.synchronize()
        res.append((float(loss), opt.flat_grad.clone(), _lib.launch_count - l0))
    (l_a, g_a, n_a), (l_b, g_b, n_b), (l_c, g_c, n_c) = res
    assert abs(l_a - l_b) <= 1e-6 * abs(l_a) and abs(l_a - l_c) <= 1e-6 * abs(l_a)
    assert relerr(g_b.cpu(), g_a.cpu()) < 1e-4 and relerr(g_c.cpu(), g_a.cpu()) < 1e-4
    assert float(g_a.abs().max()) > 0
    assert n_b == n_c and n_b < n_a - 150, (n_a, n_b, n_c)      # ~140 pack + ~75 unpack launches became two


def test_pack_table_follows_parameter_storage():
    """A model that trained with torch.optim.Adam (parameters in their own storages) and is then handed to FusedAdam (which
    re-points every parameter into one flat buffer) must rebuild its pack table: the step after the switch uses the CURRENT
    weights, not the old storages."""
    from rcfd import optim, synth
    cfg = synth.SMALL_FUSIONNET
    p0 = synth_fusionnet_state(cfg, 9)
    n, h, w = 2, 64, 96
    image, depth = synth.fusionnet_inputs(n, h, w, 9, 'quasi_dense')
    gt, lidar = synth.training_targets(n, h, w, 9)
    image, depth, gt, lidar = [t.to(DEV) for t in (image, depth, gt, lidar)]

    def step(m):
        d = m.forward(image, depth)
        loss, _ = m.compute_loss(image, d, gt, lidar, 'l1', 0.0, -1, None, 2.0)
        loss.backward()
        return float(loss)

    m = make_model(cfg, p0, precision='bf16')
    m.train()
    adam = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(2):                          # records the plan, builds the table, takes real steps
        adam.zero_grad()
        step(m)
        adam.step()
    fused = optim.FusedAdam(m.parameters(), lr=1e-2)          # parameters move into the flat buffer
    with torch.no_grad():
        for q in m.parameters():
            q.mul_(0.5)                          # and change there: the old storages keep the old values
    from rcfd import engine
    engine.note_params_changed()
    got = step(m)
    ref_model = make_model(cfg, p0, precision='bf16')
    ref_model.train()
    ref_model.encoder.load_state_dict(m.encoder.state_dict())
    ref_model.decoder.load_state_dict(m.decoder.state_dict())
    want = step(ref_model)
    assert abs(got - want) < 1e-4 * abs(want), (got, want)
