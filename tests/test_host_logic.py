"""CPU: host-side logic (engine graph wiring, reverse-mode tape, gradient delivery, checkpoint
surface, data-parallel gradient sync) with the kernels replaced by tests/mock_ops.py, a torch-CPU
emulation of the C-ABI semantics.  The real kernels are tested on the GPU (-m gpu)."""
import os
import sys

import pytest
import torch

import fusionnet_oracle as fo
from rcfd import synth
from helpers import load_golden, relerr, synth_fusionnet_state
import mock_ops


@pytest.fixture
def mocked(monkeypatch):
    import rcfd
    from rcfd import engine
    import fusionnet_model, fusionnet_losses, net_utils
    monkeypatch.setitem(sys.modules, 'rcfd.ops', mock_ops)
    monkeypatch.setattr(rcfd, 'ops', mock_ops, raising=False)
    for mod in (engine, fusionnet_model, fusionnet_losses):
        monkeypatch.setattr(mod, 'ops', mock_ops)
    return mock_ops


def _model(cfg, p):
    import fusionnet_model
    m = fusionnet_model.FusionNetModel(device=torch.device('cpu'), **cfg)
    m.encoder.load_state_dict({k[8:]: v for k, v in p.items() if k.startswith('encoder.')})
    m.decoder.load_state_dict({k[8:]: v for k, v in p.items() if k.startswith('decoder.')})
    return m


def test_engine_eval_graph_matches_golden(mocked):
    g = load_golden('fusionnet_small_2x64x96')
    p = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w::')}
    n, h, w, seed = [int(v) for v in g['meta']]
    image, depth = synth.fusionnet_inputs(n, h, w, seed, str(g['variant']))
    m = _model(synth.SMALL_FUSIONNET, p)
    m.eval()
    with torch.no_grad():
        d = m.forward(image, depth)
        l = m.forward(image, depth, return_logits=True)
    assert relerr(d, g['eval_depth']) < 1e-4 and relerr(l, g['eval_logits']) < 1e-4


def test_engine_train_step_tape_matches_golden(mocked):
    g = load_golden('fusionnet_small_2x64x96')
    p = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w::')}
    n, h, w, seed = [int(v) for v in g['meta']]
    image, depth = synth.fusionnet_inputs(n, h, w, seed, str(g['variant']))
    gt, lidar = synth.training_targets(n, h, w, seed)
    import net_utils
    m = _model(synth.SMALL_FUSIONNET, p)
    m.train()
    gt = net_utils.OutlierRemoval(7, 1.5).remove_outliers(gt)
    d = m.forward(image, depth)
    # is_cuda is False on the mock: the tensor-op branch of compute_loss is exercised here ...
    loss, info = m.compute_loss(image=image, output_depth=d, ground_truth=gt, lidar_map=lidar, loss_func='l1',
                                w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                                validity_map_loss_smoothness=torch.ones_like(gt), w_lidar_loss=2.0)
    loss.backward()
    assert relerr(d.detach(), g['train_depth']) < 1e-4
    assert abs(float(loss) - float(g['train_loss'])) < 1e-4 * float(g['train_loss'])
    named = dict([('encoder.' + k, v) for k, v in m.encoder.named_parameters()] +
                 [('decoder.' + k, v) for k, v in m.decoder.named_parameters()])
    for i, k in enumerate(g['grad_names']):
        k = str(k)
        if g['grad_none'][i]:
            assert named[k].grad is None, k
        else:
            assert abs(float(named[k].grad.double().sum()) - g['grad_sum'][i]) <= 1e-3 * g['grad_abs'][i] + 1e-9, k
            assert relerr(named[k].grad.flatten()[:8], g['grad_head'][i][:named[k].numel()]) < 5e-3 or \
                float(named[k].grad.abs().max()) < 1e-6, k
    sd = m.decoder.state_dict()
    assert relerr(sd['deconv0.conv.batch_norm.running_mean'], g['train_running_mean_deconv0']) < 1e-4
    assert int(sd['deconv0.conv.batch_norm.num_batches_tracked']) == 1
    # ... and the fused masked-L1 autograd node gives the same loss / gradient
    import fusionnet_losses
    d2 = d.detach().clone().requires_grad_(True)
    l2 = fusionnet_losses.MaskedL1.apply(d2, gt, lidar, 2.0)
    l2.backward()
    d3 = d.detach().clone().requires_grad_(True)
    fo.fusionnet_loss(d3, gt, lidar, 2.0, 'l1').backward()
    assert abs(float(l2) - float(loss)) < 1e-6 * float(loss) and relerr(d2.grad, d3.grad) < 1e-6


def test_engine_multires_decoder_matches_golden(mocked):
    """n_resolution_decoder = 3 through the engine (kernels mocked): outputs at every scale, multi-scale loss, every
    gradient vs the fixture written by the reference; the graph wiring (logit outputs, bilinear up-sampling, padded concat
    behind the skips, gradient accumulation on the logits from both consumers) is what is under test here."""
    from helpers import multires_case_inputs, check_multires_against_golden, fusionnet_state_template
    g, cfg, seed, image, depth, weights = multires_case_inputs()
    p = fusionnet_state_template(cfg)
    synth.fill_state_dict_(p, seed)
    m = _model(cfg, p)
    m.train()
    outs = m.forward(image, depth, return_multiscale=True)
    loss = sum(wi * o.mean() for wi, o in zip(weights, outs))
    loss.backward()
    named = dict([('encoder.' + k, v) for k, v in m.encoder.named_parameters()] +
                 [('decoder.' + k, v) for k, v in m.decoder.named_parameters()])
    check_multires_against_golden(g, outs, named, loss)
    assert torch.equal(m.forward(image, depth).detach() * 0, outs[-1].detach() * 0)       # single output: the last scale
    with pytest.raises(ValueError):
        import networks
        networks.MultiScaleDecoder(n_resolution=5, n_filters=[32, 32, 32, 16, 16, 16], n_skips=[32, 32, 16, 16, 16, 0])


def test_tensor_core_parity_mode_host_logic(mocked):
    """set_precision('bf16x3' | 'bf16x6'): fp32 storage, every conv / weight gradient = 3 / 6 bf16-operand passes over
    2- / 3-part splits (rcfd/x3.py, the REAL orchestration; the passes themselves are emulated with bf16-rounded operands
    and fp32 accumulation, which is what tcgen05 kind::f16 computes).  Checks the wiring (packed (hi, lo) weight pairs, the
    space-to-depth stems, sub-pixel weights, statistics / epilogue after the third pass, x3 weight gradients) and the
    numerical claim: the split reproduces the reference's fp32 results far inside the 1e-3 contract, while ONE bf16
    pass (what the fast mode computes) does not."""
    g = load_golden('fusionnet_small_2x64x96')
    p = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('w::')}
    n, h, w, seed = [int(v) for v in g['meta']]
    image, depth = synth.fusionnet_inputs(n, h, w, seed, str(g['variant']))
    gt, lidar = synth.training_targets(n, h, w, seed)
    m = _model(synth.SMALL_FUSIONNET, p)
    m.set_precision('bf16x3')
    m.eval()
    with torch.no_grad():
        l = m.forward(image, depth, return_logits=True)
        d = m.forward(image, depth)
    assert relerr(l, g['eval_logits']) < 1e-4 and relerr(d, g['eval_depth']) < 1e-4
    import net_utils
    m.set_precision('bf16x6')        # 3-part split, 6 passes: fp32-class, what the gradient parity tests use
    m.train()
    gt = net_utils.OutlierRemoval(7, 1.5).remove_outliers(gt)
    d = m.forward(image, depth)
    loss, _ = m.compute_loss(image, d, gt, lidar, 'l1', 0.0, -1, torch.ones_like(gt), 2.0)
    loss.backward()
    assert relerr(d.detach(), g['train_depth']) < 1e-5
    assert abs(float(loss) - float(g['train_loss'])) < 1e-4 * float(g['train_loss'])
    named = dict([('encoder.' + k, v) for k, v in m.encoder.named_parameters()] +
                 [('decoder.' + k, v) for k, v in m.decoder.named_parameters()])
    for i, k in enumerate(g['grad_names']):
        k = str(k)
        if g['grad_none'][i]:
            assert named[k].grad is None, k
        else:
            assert abs(float(named[k].grad.double().sum()) - g['grad_sum'][i]) <= 1e-3 * g['grad_abs'][i] + 1e-9, k
    # one bf16 pass is NOT inside the contract (the reason this mode exists): hi-only weights and activations
    hi = lambda t: t.to(torch.bfloat16).float()
    with torch.no_grad():
        x = torch.randn(1, 8, 8, 64)
        wt = torch.randn(32, 9, 64) * 0.05
        full = mock_ops.conv2d(x, wt, 32, 3)
        one = mock_ops.conv2d(hi(x), hi(wt), 32, 3)
        three = mock_ops.conv2d(x, mock_ops.split_bf16(wt, 2), 32, 3)
        six = mock_ops.conv2d(x, mock_ops.split_bf16(wt, 3), 32, 3)
    assert relerr(six, full) < 2e-6 < relerr(three, full) < 2e-5 < 1e-3 < relerr(one, full)


def test_checkpoint_surface_roundtrip(tmp_path, mocked):
    p = synth_fusionnet_state(synth.SMALL_FUSIONNET, 7)
    m = _model(synth.SMALL_FUSIONNET, p)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    path = str(tmp_path / 'model-5.pth')
    m.save_model(path, 5, opt)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {'train_step', 'optimizer_state_dict', 'encoder_state_dict', 'decoder_state_dict'}
    # written like the reference does after data_parallel(): 'module.'-prefixed keys
    assert all(k.startswith('module.') for part in ('encoder_state_dict', 'decoder_state_dict') for k in ck[part])
    m3 = _model(synth.SMALL_FUSIONNET, synth_fusionnet_state(synth.SMALL_FUSIONNET, 9))
    assert m3.restore_model(path)[0] == 5
    # bare keys (a checkpoint saved without DataParallel) must load too
    ck['encoder_state_dict'] = {k[len('module.'):]: v for k, v in ck['encoder_state_dict'].items()}
    ck['decoder_state_dict'] = {k[len('module.'):]: v for k, v in ck['decoder_state_dict'].items()}
    torch.save(ck, path)
    m2 = _model(synth.SMALL_FUSIONNET, synth_fusionnet_state(synth.SMALL_FUSIONNET, 8))
    step, _ = m2.restore_model(path, torch.optim.Adam(m2.parameters(), lr=1e-3))
    assert step == 5
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        import fusionnet_model
        fusionnet_model.FusionNetModel(device=torch.device('cpu'), **dict(synth.SMALL_FUSIONNET, fusion_type='nope'))
    with pytest.raises(ValueError):
        m.compute_loss(None, torch.ones(1, 1, 2, 2), torch.ones(1, 1, 2, 2), torch.ones(1, 1, 2, 2), 'bogus', 0.0, -1,
                       None, 0.0)


def test_reference_checkpoint_interop(tmp_path, mocked):
    """tests/golden/reference_checkpoint_tiny.pth was written by the UNMODIFIED reference (DataParallel-wrapped model,
    torch.optim.Adam after one step; tests/golden/make_golden.py checkpoint_case, which also checks on the spot that the
    reference loads a checkpoint written here).  It restores into the product model + FusedAdam, the round trip through
    the product's save_model reproduces the reference file's structure ('module.' keys, torch.optim.Adam layout), and
    FusedAdam <-> torch.optim.Adam state dicts interchange."""
    from rcfd import optim
    import fusionnet_model
    cfg = dict(synth.CANONICAL_FUSIONNET, n_filters_encoder_image=[8, 8, 16, 16, 16, 16],
               n_filters_encoder_depth=[8, 8, 8, 8, 8, 8], n_filters_decoder=[16, 16, 16, 8, 8, 8])
    ref_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_checkpoint_tiny.pth')
    ref_ck = torch.load(ref_path, weights_only=False)
    m = fusionnet_model.FusionNetModel(device=torch.device('cpu'), **cfg)
    opt = optim.FusedAdam([{'params': m.parameters(), 'weight_decay': 0.0}], lr=5e-4)
    step, opt = m.restore_model(ref_path, optimizer=opt)
    assert step == 7 and opt.step_count == 1
    ref_state = ref_ck['optimizer_state_dict']['state']
    params = m.parameters()
    off = 0
    for i, p in enumerate(params):
        n = p.numel()
        if i in ref_state:
            assert torch.equal(opt.exp_avg[off:off + n].view(p.shape), ref_state[i]['exp_avg'])
            assert torch.equal(opt.exp_avg_sq[off:off + n].view(p.shape), ref_state[i]['exp_avg_sq'])
        else:                                    # never-used projection weights: torch's Adam has no state for them
            assert not bool(opt.exp_avg[off:off + n].any())
        assert torch.equal(p, ref_ck['encoder_state_dict' if i < len(list(m.encoder.parameters())) else 'decoder_state_dict'][
            'module.' + dict((id(q), k) for root in (m.encoder, m.decoder) for k, q in root.named_parameters())[id(p)]])
        off += n
    assert len(params) - 14 <= len(ref_state) < len(params)      # identity-shortcut projections never get Adam state
    # product save -> same structure as the reference's file
    out = str(tmp_path / 'model-9.pth')
    m.save_model(out, 9, opt)
    ck = torch.load(out, weights_only=False)
    assert set(ck) == set(ref_ck)
    for part in ('encoder_state_dict', 'decoder_state_dict'):
        assert list(ck[part]) == list(ref_ck[part])
        for k in ck[part]:
            assert ck[part][k].shape == ref_ck[part][k].shape and torch.equal(ck[part][k], ref_ck[part][k]), k
    sd, rsd = ck['optimizer_state_dict'], ref_ck['optimizer_state_dict']
    assert set(sd['state']) == set(rsd['state'])
    assert set(rsd['param_groups'][0]) <= set(sd['param_groups'][0])
    assert sd['param_groups'][0]['params'] == rsd['param_groups'][0]['params']
    # torch.optim.Adam accepts FusedAdam's state and vice versa
    plain = [torch.nn.Parameter(p.detach().clone()) for p in params]
    adam = torch.optim.Adam([{'params': plain, 'weight_decay': 0.0}], lr=1e-3)
    adam.load_state_dict(sd)
    assert torch.equal(adam.state_dict()['state'][0]['exp_avg'], ref_state[0]['exp_avg'])
    m2 = fusionnet_model.FusionNetModel(device=torch.device('cpu'), **cfg)
    opt2 = optim.FusedAdam(m2.parameters(), lr=1e-3)
    opt2.load_state_dict(adam.state_dict())
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq) and opt2.step_count == 1
    # a fresh FusedAdam (no step yet) has an empty per-parameter state, like torch's
    assert optim.FusedAdam(fusionnet_model.FusionNetModel(device=torch.device('cpu'), **cfg).parameters()).state_dict()['state'] == {}


def _ddp_worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for q in (here, os.path.join(root, 'oracle'), os.path.join(root, 'radar-camera-fusion-depth_b200')):
        sys.path.insert(0, q)
    import rcfd
    from rcfd import engine, parallel
    import fusionnet_model, fusionnet_losses
    sys.modules['rcfd.ops'] = mock_ops
    rcfd.ops = mock_ops
    for mod in (engine, fusionnet_model, fusionnet_losses):
        mod.ops = mock_ops
    torch.manual_seed(100 + rank)                       # different init per rank: broadcast must fix it
    cfg = dict(synth.SMALL_FUSIONNET)
    m = fusionnet_model.FusionNetModel(device=torch.device('cpu'), **cfg)
    m.data_parallel()                                   # attaches DistributedGradSync, broadcasts rank 0's state
    assert isinstance(m.grad_hook, parallel.DistributedGradSync)
    m.train()
    image, depth = synth.fusionnet_inputs(1, 64, 64, 50 + rank, 'quasi_dense')
    gt, lidar = synth.training_targets(1, 64, 64, 50 + rank)
    d = m.forward(image, depth)
    loss, _ = m.compute_loss(image, d, gt, lidar, 'l1', 0.0, -1, torch.ones_like(gt), 2.0)
    loss.backward()
    # the overlapped form: slices of a flat buffer reduced asynchronously (tail first), finish() waits and averages
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    m.grad_hook.flat_grad = flat
    m.grad_hook.reduce_slice(6, None)
    m.grad_hook.reduce_slice(0, 6)
    m.grad_hook.finish()
    torch.save({'grads': [None if q.grad is None else q.grad.clone() for q in m.parameters()],
                'params': [q.detach().clone() for q in m.parameters()], 'sliced': flat.clone(),
                'payload': m.grad_hook.payload_bytes()}, os.path.join(tmp, 'rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_sync_gloo(tmp_path):
    """world_size 2 over gloo: replicas start identical (rank-0 broadcast), every used gradient is
    the average over ranks, never-used projections stay None and outside the buckets."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(str(tmp_path / 'rank0.pt'), weights_only=False)
    r1 = torch.load(str(tmp_path / 'rank1.pt'), weights_only=False)
    n_none = 0
    for a, b in zip(r0['params'], r1['params']):
        assert torch.equal(a, b)
    for a, b in zip(r0['grads'], r1['grads']):
        assert (a is None) == (b is None)
        if a is None:
            n_none += 1
        else:
            assert torch.equal(a, b)
    # SMALL config: the 10 identity-shortcut blocks '*.1.projection' plus blocks2_image.0 (16 -> 16, stride 1)
    assert n_none == 11
    used = sum(p.numel() for p, g in zip(r0['params'], r0['grads']) if g is not None)
    assert r0['payload'] == used * 4
    assert torch.equal(r0['sliced'], torch.arange(10, dtype=torch.float32) * 1.5) and torch.equal(r0['sliced'], r1['sliced'])


def test_two_part_backward_equals_whole_tape(mocked):
    """Tape.backward(part=0) + backward(part=1) around the split marker (encoder level 5) == one backward(): the same
    parameter gradients (the data-parallel step captures the two parts as two CUDA graphs and all-reduces the gradients
    of the first part while the second runs)."""
    import fusionnet_model
    from rcfd import engine
    cfg = dict(synth.SMALL_FUSIONNET)
    torch.manual_seed(0)
    m = fusionnet_model.FusionNetModel(device=torch.device('cpu'), **cfg)
    m.train()
    image, depth = synth.fusionnet_inputs(1, 64, 64, 3, 'quasi_dense')
    results = []
    for two_part in (False, True):
        out, ectx = m._run(image, depth, record=True)
        tape = ectx.tape
        assert tape.split_at is not None and 0 < tape.split_at < len(tape.steps)
        tape.set_grad(out, torch.ones_like(out))
        if two_part:
            tape.backward(part=0)
            n_first = len(tape.param_grads)
            assert 0 < n_first and tape.steps            # part 0 delivered the decoder / level >= 5 gradients only
            tape.backward(part=1)
            assert len(tape.param_grads) > n_first and not tape.steps
        else:
            tape.backward()
        results.append({id(p): g.clone() for p, g in tape.param_grads})
    assert set(results[0]) == set(results[1])
    for k in results[0]:
        assert torch.allclose(results[0][k], results[1][k], rtol=1e-5, atol=1e-7)


def test_batched_transforms_match_reference_fixture():
    """fusionnet_transforms.Transforms (batched tensor expressions, same constructor / transform surface as the
    reference) reproduces, bit for bit, what the reference's per-sample loop produced on the same seeded inputs and
    random draws (tests/golden/transforms_5x18x26.npz, written by the reference)."""
    import numpy as np
    import fusionnet_transforms
    from helpers import load_golden
    g = load_golden('transforms_5x18x26')
    seed = int(g['meta'][0])
    cfg = dict(random_brightness=[0.8, 1.2], random_contrast=[0.8, 1.2], random_saturation=[0.8, 1.2],
               random_flip_type=['horizontal', 'vertical'])
    for tag, scale, rng in (('u8_01', 255.0, [0, 1]), ('f_pm1', 1.0, [-1, 1])):
        gen = torch.Generator().manual_seed(seed)
        img = torch.rand(5, 3, 18, 26, generator=gen) * scale
        if scale > 1.0:
            img = img.round()
        maps = [torch.rand(5, 1, 18, 26, generator=gen) * 50, torch.rand(5, 2, 18, 26, generator=gen)]
        t = fusionnet_transforms.Transforms(normalized_image_range=rng, **cfg)
        torch.manual_seed(seed + 1)
        (oi,), om = t.transform([img.clone()], [m.clone() for m in maps], random_transform_probability=0.9)
        assert np.array_equal(oi.numpy(), g[tag + '_image'])
        assert np.array_equal(om[0].numpy(), g[tag + '_map0']) and np.array_equal(om[1].numpy(), g[tag + '_map1'])
    # images only -> a bare list, like the reference
    t = fusionnet_transforms.Transforms(normalized_image_range=[0, 1])
    out = t.transform([torch.rand(2, 3, 4, 4) * 255], random_transform_probability=0.0)
    assert isinstance(out, list) and out[0].shape == (2, 3, 4, 4) and float(out[0].max()) <= 1.0


def test_losses_and_compute_loss_branches_match_reference_fixture():
    """fusionnet_losses.* and the non-canonical branches of FusionNetModel.compute_loss (l2 / smoothl1, first-order and
    Sobel smoothness, with / without the lidar term; tensor-op formulas) against values computed by the reference
    (tests/golden/losses_2x24x40.npz)."""
    import numpy as np
    import fusionnet_losses as L
    import fusionnet_model
    from helpers import load_golden, relerr
    g = load_golden('losses_2x24x40')
    seed, n, h, w = [int(v) for v in g['meta']]
    gen = torch.Generator().manual_seed(seed)
    image = torch.rand(n, 3, h, w, generator=gen)
    pred = torch.rand(n, 1, h, w, generator=gen) * 50 + 1
    tgt = torch.rand(n, 1, h, w, generator=gen) * 50 + 1
    weights = (torch.rand(n, 1, h, w, generator=gen) > 0.3).float()
    gt = tgt * (torch.rand(n, 1, h, w, generator=gen) < 0.4)
    lidar = (torch.rand(n, 1, h, w, generator=gen) * 50 + 1) * (torch.rand(n, 1, h, w, generator=gen) < 0.05)
    tol = 1e-6
    assert relerr(L.l1_loss(pred, tgt), g['l1']) < tol
    assert relerr(L.l2_loss(pred, tgt), g['l2']) < tol
    assert relerr(L.smooth_l1_loss(pred, tgt), g['smoothl1']) < tol
    assert relerr(L.smoothness_loss_func(pred, image), g['smooth']) < tol
    assert relerr(L.sobel_smoothness_loss_func(pred, image, weights, [1, 1, 7, 7]), g['sobel7']) < tol
    assert relerr(L.sobel_smoothness_loss_func(pred, image, weights, [1, 1, 3, 3]), g['sobel3']) < tol
    gx, gy = L.sobel_filter([1, 1, 7, 7])
    assert np.array_equal(gx.numpy(), g['sobel_gx']) and np.array_equal(gy.numpy(), g['sobel_gy'])
    dy, dx = L.gradient_yx(pred)
    assert np.array_equal(dy.numpy(), g['grad_dy']) and np.array_equal(dx.numpy(), g['grad_dx'])
    # compute_loss on CPU tensors takes the tensor-op branches (the fused kernel needs CUDA tensors)
    m = fusionnet_model.FusionNetModel.__new__(fusionnet_model.FusionNetModel)
    combos = [('l1', 0.0, -1, 2.0), ('l2', 0.0, -1, 2.0), ('smoothl1', 0.0, -1, 0.0), ('l1', 0.5, -1, 2.0),
              ('l1', 0.5, 7, 2.0), ('l2', 0.25, 3, 0.0)]
    for (lf, ws, ks, wl), ref in zip(combos, g['compute_loss']):
        loss, info = m.compute_loss(image=image, output_depth=pred, ground_truth=gt, lidar_map=lidar, loss_func=lf,
                                    w_smoothness=ws, loss_smoothness_kernel_size=ks,
                                    validity_map_loss_smoothness=weights, w_lidar_loss=wl)
        assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref)), (lf, ws, ks, wl, float(loss), float(ref))
    with pytest.raises(ValueError):
        m.compute_loss(image, pred, gt, lidar, 'huber', 0.0, -1, weights, 0.0)


class _DummyDepthModel(object):
    def forward(self, image, input_depth):
        return 1.0 + 40.0 * image.mean(dim=1, keepdim=True) + 0.5 * input_depth[:, 0:1]


def test_validate_matches_reference_fixture(tmp_path):
    """fusionnet_main.validate + eval_utils: the reference's metrics (mm, 1/km), masks and best-result rule, against
    numbers and the log text produced by the reference's own validate() on the same seeded samples and dummy model
    (tests/golden/validate_3x20x32.npz)."""
    import numpy as np
    import fusionnet_main
    import fusionnet_transforms
    from helpers import load_golden
    g = load_golden('validate_3x20x32')
    gen = torch.Generator().manual_seed(int(g['meta'][0]))
    loader = []
    for _ in range(3):
        image = (torch.rand(1, 3, 20, 32, generator=gen) * 255).round()
        depth = torch.rand(1, 1, 20, 32, generator=gen) * 60 * (torch.rand(1, 1, 20, 32, generator=gen) < 0.2)
        response = torch.rand(1, 1, 20, 32, generator=gen)
        gt = torch.rand(1, 1, 20, 32, generator=gen) * 90 * (torch.rand(1, 1, 20, 32, generator=gen) < 0.5)
        loader.append((image, depth, response, gt))
    tr = fusionnet_transforms.Transforms(normalized_image_range=[0, 1])
    log_path = str(tmp_path / 'log.txt')
    best = {'step': -1, 'mae': np.inf, 'rmse': np.inf, 'imae': np.inf, 'irmse': np.inf}
    best = fusionnet_main.validate(_DummyDepthModel(), loader, tr, step=10, best_results=best, min_evaluate_depth=1.0,
                                   max_evaluate_depth=80.0, device=torch.device('cpu'), summary_writer=None, log_path=log_path)
    first = np.array([best['step'], best['mae'], best['rmse'], best['imae'], best['irmse']], dtype=np.float64)
    assert np.allclose(first, g['first'], rtol=1e-9, atol=0)
    worse = dict(best)
    worse.update(mae=best['mae'] - 1.0, rmse=best['rmse'] - 1.0)
    second = fusionnet_main.validate(_DummyDepthModel(), loader, tr, step=20, best_results=dict(worse), min_evaluate_depth=1.0,
                                     max_evaluate_depth=80.0, device=torch.device('cpu'), summary_writer=None, log_path=log_path)
    got = np.array([second['step'], second['mae'], second['rmse'], second['imae'], second['irmse']], dtype=np.float64)
    assert np.allclose(got, g['second'], rtol=1e-9, atol=0) and int(second['step']) == 10
    assert open(log_path).read() == str(g['log'])


def test_train_cli_flag_surface_matches_reference():
    """train_fusionnet.py exposes every flag of the reference's CLI (src/train_fusionnet.py:8-130; list taken from
    the reference) plus --precision / --max_steps, and train() accepts every resulting keyword."""
    import inspect
    import fusionnet_main
    import train_fusionnet
    reference_flags = [
        'activation_func', 'augmentation_probabilities', 'augmentation_random_brightness', 'augmentation_random_contrast',
        'augmentation_random_crop_type', 'augmentation_random_flip_type', 'augmentation_random_saturation',
        'augmentation_schedule', 'batch_size', 'checkpoint_dirpath', 'decoder_type', 'device', 'encoder_type', 'fusion_type',
        'ground_truth_dilation_kernel_size', 'input_channels_depth', 'input_channels_image', 'learning_rates',
        'learning_schedule', 'loss_func', 'loss_smoothness_kernel_size', 'max_evaluate_depth', 'max_predict_depth',
        'min_evaluate_depth', 'min_predict_depth', 'n_filters_decoder', 'n_filters_encoder_depth', 'n_filters_encoder_image',
        'n_height', 'n_resolutions_decoder', 'n_step_per_checkpoint', 'n_step_per_summary', 'n_thread', 'n_width',
        'normalized_image_range', 'outlier_removal_kernel_size', 'outlier_removal_threshold', 'restore_path',
        'start_step_validation', 'train_depth_path', 'train_ground_truth_path', 'train_image_path', 'train_lidar_map_path',
        'train_response_path', 'val_depth_path', 'val_ground_truth_path', 'val_image_path', 'val_response_path',
        'w_lidar_loss', 'w_smoothness', 'w_weight_decay', 'weight_initializer']
    mine = {a.dest for a in train_fusionnet.parser._actions if a.dest != 'help'}
    assert set(reference_flags) <= mine
    assert mine - set(reference_flags) == {'precision', 'max_steps'}
    args = vars(train_fusionnet.parser.parse_args([]))
    args['ground_truth_outlier_removal_kernel_size'] = args.pop('outlier_removal_kernel_size')
    args['ground_truth_outlier_removal_threshold'] = args.pop('outlier_removal_threshold')
    params = inspect.signature(fusionnet_main.train).parameters
    assert set(args) <= set(params)
    required = {k for k, p in params.items() if p.default is inspect.Parameter.empty}
    assert required <= set(args)                       # the CLI supplies every required keyword of train()


def test_run_cli_flag_surface_matches_reference():
    """run_fusionnet.py exposes every flag of the reference's CLI (src/run_fusionnet.py:8-67; list taken from the
    reference) plus --precision, and run() accepts every resulting keyword in the reference's order."""
    import inspect
    import fusionnet_main
    import run_fusionnet
    reference_flags = [
        'restore_path', 'image_path', 'depth_path', 'response_path', 'ground_truth_path', 'input_channels_image',
        'input_channels_depth', 'normalized_image_range', 'encoder_type', 'n_filters_encoder_image',
        'n_filters_encoder_depth', 'fusion_type', 'decoder_type', 'n_filters_decoder', 'n_resolutions_decoder',
        'min_predict_depth', 'max_predict_depth', 'weight_initializer', 'activation_func', 'output_dirpath', 'save_outputs',
        'keep_input_filenames', 'verbose', 'min_evaluate_depth', 'max_evaluate_depth']
    mine = [a.dest for a in run_fusionnet.parser._actions if a.dest != 'help']
    assert mine == reference_flags + ['precision']
    params = list(inspect.signature(fusionnet_main.run).parameters)
    assert params == reference_flags + ['precision']         # the reference passes everything by keyword; same names, same order
    required = {a.dest for a in run_fusionnet.parser._actions if a.required}
    assert required == {'restore_path', 'image_path', 'depth_path', 'response_path', 'output_dirpath'}


def test_png16_writer_matches_reference_codec(tmp_path):
    """rcfd.data.save_png16 == the reference's save_depth / save_response (uint32(z * multiplier) as a mode-'I' PNG):
    the file decodes to what the reference's loaders return."""
    import numpy as np
    from PIL import Image
    from rcfd import data
    z = (np.random.RandomState(0).rand(12, 20).astype(np.float32) * 90) * (np.random.RandomState(1).rand(12, 20) < 0.5)
    data.save_png16(z, str(tmp_path / 'd.png'), data.DEPTH_MULTIPLIER)
    back = np.array(Image.open(str(tmp_path / 'd.png')), dtype=np.float32) / 256.0        # load_depth (src/data_utils.py:254-257)
    assert np.array_equal(back, np.floor(z * 256.0) / 256.0)
    assert np.array_equal(data.load_png16(str(tmp_path / 'd.png')), np.uint16(np.uint32(z * 256.0)))


def test_radarnet_compute_loss_matches_reference_fixture():
    """RadarNetModel.compute_loss (weighted BCE over valid pixels, reference src/radarnet_model.py:126-167) against
    values computed by the reference."""
    import radarnet_model
    from helpers import load_golden
    g = load_golden('radarnet_loss_3x64x64')
    gen = torch.Generator().manual_seed(int(g['meta'][0]))
    logits = torch.randn(3, 1, 64, 64, generator=gen) * 3
    gt = (torch.rand(3, 1, 64, 64, generator=gen) < 0.2).float()
    valid = (torch.rand(3, 1, 64, 64, generator=gen) < 0.7).float()
    m = radarnet_model.RadarNetModel.__new__(radarnet_model.RadarNetModel)
    for w, ref in zip((1.0, 2.0, 5.5), g['loss']):
        loss, info = m.compute_loss(logits, gt, valid, w_positive_class=w)
        assert abs(float(loss) - float(ref)) < 1e-6 * abs(float(ref)) and 'loss' in info


def test_data_path_host_side(tmp_path):
    """rcfd.data without the reference on sys.path: path lists, 16-bit PNG rasters (written here with Pillow the way the
    reference's save_depth does, src/data_utils.py:271-286), the raw dataset's on-disk sample types, and crop_origin's
    draws against the decisions of the reference's random_crop (src/datasets.py:19-109) replayed from the same seed."""
    import numpy as np
    from PIL import Image
    from rcfd import data
    rng = np.random.RandomState(3)
    h0, w0 = 40, 70
    files = {k: [] for k in ('image', 'depth', 'response', 'gt', 'lidar')}
    truth = []
    for i in range(3):
        img = rng.randint(0, 256, (h0, w0, 3)).astype(np.uint8)
        Image.fromarray(img).save(str(tmp_path / ('im%d.png' % i)))
        files['image'].append(str(tmp_path / ('im%d.png' % i)))
        maps = {}
        for k in ('depth', 'response', 'gt', 'lidar'):
            z = (rng.rand(h0, w0) * 80 * (rng.rand(h0, w0) < 0.4)).astype(np.float32)
            Image.fromarray(np.uint32(z * 256.0), mode='I').save(str(tmp_path / ('%s%d.png' % (k, i))))     # save_depth
            files[k].append(str(tmp_path / ('%s%d.png' % (k, i))))
            maps[k] = z
        truth.append((img, maps))
    for k, v in files.items():
        with open(str(tmp_path / (k + '.txt')), 'w') as f:
            f.write('\n'.join(v) + '\n')
    paths = {k: data.read_paths(str(tmp_path / (k + '.txt'))) for k in files}
    assert paths['image'] == files['image'] and len(paths['gt']) == 3
    z16 = data.load_png16(files['depth'][1])
    assert z16.dtype == np.uint16 and np.array_equal(z16, np.uint32(truth[1][1]['depth'] * 256.0).astype(np.uint16))
    # load_depth of the reference == raster / 256 with <= 0 -> 0
    ref = np.array(Image.open(files['depth'][1]), dtype=np.float32) / 256.0
    ref[ref <= 0] = 0.0
    assert np.array_equal(z16.astype(np.float32) / 256.0, ref)
    ds = data.FusionNetRawDataset(paths['image'], paths['depth'], paths['response'], paths['gt'], paths['lidar'],
                                  shape=(24, 32), random_crop_type=['horizontal', 'vertical'])
    np.random.seed(11)
    sample = ds[2]
    assert sample[0].dtype == torch.uint8 and tuple(sample[0].shape) == (24, 32, 3)
    assert sample[1].dtype == torch.int16 and tuple(sample[1].shape) == (24, 32) and tuple(sample[5]) == (0, 0)
    np.random.seed(11)
    y0, x0 = data.crop_origin(h0, w0, 24, 32, ['horizontal', 'vertical'])
    assert np.array_equal(sample[0].numpy(), truth[2][0][y0:y0 + 24, x0:x0 + 32])
    full = data.FusionNetRawDataset(paths['image'], paths['depth'], paths['response'], paths['gt'], paths['lidar'],
                                    shape=(24, 32), random_crop_type=['horizontal', 'vertical'], crop_on_host=False)
    np.random.seed(11)
    s2 = full[2]
    assert tuple(s2[0].shape) == (h0, w0, 3) and tuple(int(v) for v in s2[5]) == (y0, x0)
    # crop decisions: origins chosen by the reference's own random_crop (tests/golden/crop_origins_90x160.npz, written by
    # make_golden.py crop_case from src/datasets.py:19-109) for every crop type under the same seeded draws
    from helpers import load_golden
    g = load_golden('crop_origins_90x160')
    oh, ow, nh, nw = [int(v) for v in g['meta']]
    crop_types = (['none'], ['center'], ['left', 'top'], ['right', 'bottom'], ['horizontal'], ['horizontal', 'anchored'],
                  ['horizontal', 'vertical'], ['horizontal', 'vertical', 'anchored'], ['bottom', 'horizontal'])
    for ci, crop_type in enumerate(crop_types):
        for seed in range(8):
            got = data.crop_origin(oh, ow, nh, nw, crop_type, rng=np.random.RandomState(100 + seed))
            assert got == tuple(int(v) for v in g['origins'][ci, seed]), (crop_type, seed)
    # synthetic batch -> file precision -> back (what bench.py's end-to-end leg ships over PCIe)
    image = (torch.rand(2, 3, 8, 12) * 255).round()
    maps = [torch.rand(2, 1, 8, 12) * 80 * (torch.rand(2, 1, 8, 12) < 0.5) for _ in range(4)]
    raw = data.encode_raw_batch(image, *maps)
    assert raw[0].dtype == torch.uint8 and tuple(raw[0].shape) == (2, 8, 12, 3) and raw[1].dtype == torch.int16
    back = raw[1].view(torch.uint16).to(torch.int32).float() / 256.0
    assert float((back - maps[0][:, 0]).abs().max()) <= 1.0 / 256.0
