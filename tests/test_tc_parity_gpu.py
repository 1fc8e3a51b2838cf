"""GPU parity of the TENSOR-CORE path against the CPU oracle at the BASELINE sizes.

The bf16 fast mode (one tcgen05 pass over bf16-rounded operands) cannot meet the north star's 1e-3 contract; the
parity modes run the SAME tcgen05 / TMA engines over bf16 splits of the fp32 operands ('bf16x3': 2 parts / 3 passes,
'bf16x6': 3 parts / 6 passes, rcfd/x3.py) and do.  Checked here, through the C-ABI:
  * every engine family, one layer each, against float64 convolution (unit level);
  * BASELINE configs[0]: forward 1 x 320 x 576 vs the oracle, <= 1e-3 (logits, depth, MAE);
  * BASELINE configs[1] shape: one full training step at 352 x 704, batch 2 -- depth, loss, EVERY parameter gradient,
    BatchNorm running statistics -- vs the oracle;
  * BASELINE configs[2] shape: RadarNet 352 x 704 (+ 2 x 144 padding), K = 64 points, patch 352 x 288 vs the oracle;
  * the deviation of the bf16 FAST mode from the oracle is measured at the same sizes, asserted at a stated bound and
    written to gpurun_out/r2_bf16_deviation.json (copied to profiles/ and quoted in DESIGN.md / bench.py).
Tolerances are stated per assertion."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import fusionnet_oracle as fo
import radarnet_oracle as ro
from rcfd import synth
from helpers import relerr, synth_fusionnet_state

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3


def record(key, values):
    """Append measured deviations to gpurun_out/r2_bf16_deviation.json (merged back by gpurun)."""
    path = os.path.join(ROOT, 'gpurun_out', 'r2_bf16_deviation.json')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = values
    json.dump(data, open(path, 'w'), indent=1, sort_keys=True)


def make_model(cfg, p, precision):
    import fusionnet_model
    m = fusionnet_model.FusionNetModel(device=DEV, **cfg)
    m.encoder.load_state_dict({k[len('encoder.'):]: v for k, v in p.items() if k.startswith('encoder.')})
    m.decoder.load_state_dict({k[len('decoder.'):]: v for k, v in p.items() if k.startswith('decoder.')})
    m.set_precision(precision)
    return m


# ----------------------------------------------------------------------------- unit level: one layer per engine family
def _ref_conv(x, w, stride=1, up=None, x1=None):
    a = x.double().permute(0, 3, 1, 2)
    if up is not None:
        a = F.interpolate(a, size=up)
    if x1 is not None:
        a = torch.cat([a, x1.double().permute(0, 3, 1, 2)], 1)
    return F.conv2d(a, w.double(), None, stride, w.shape[-1] // 2).permute(0, 2, 3, 1)


CASES = [
    # name, n, h, w, c0, c1, cout, k, stride, up2x     (engine AUTO picks: comment)
    ('strip_is_64_64', 2, 96, 160, 64, 0, 64, 3, 1, False),        # row-streaming, input-stationary
    ('strip_concat_64_32', 1, 96, 256, 64, 32, 64, 3, 1, False),   # row-streaming over torch.cat
    ('strip_up_64_32', 1, 64, 128, 64, 0, 32, 3, 1, True),         # sub-pixel up-conv, row-streaming
    ('tma_256_256', 2, 22, 44, 256, 0, 256, 3, 1, False),          # per-tap TMA engine, paired k-steps
    ('tma_s2_64_128', 2, 44, 88, 64, 0, 128, 3, 2, False),         # per-tap TMA engine, stride 2
    ('tma_1x1_128_512', 2, 22, 44, 128, 0, 512, 1, 1, False),      # stacked 1x1 fusion GEMM
    ('tma_concat_256_256', 1, 22, 44, 256, 256, 256, 3, 1, False),  # decoder conv over concat
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize('parts', [2, 3])
def test_conv_and_wgrad_split_passes_vs_float64(case, parts):
    """conv + weight gradient through the tcgen05 engines in the parity modes vs float64: 2 parts <= 5e-5, 3 parts <= 1e-5
    (relative to the output's max; fp32 accumulation over K <= 4608 terms)."""
    from rcfd import ops
    name, n, h, w, c0, c1, cout, k, stride, up = case
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, h, w, c0, generator=g)
    x1 = torch.randn(n, 2 * h if up else h, 2 * w if up else w, c1, generator=g) if c1 else None
    wt = torch.randn(cout, c0 + c1, k, k, generator=g) / (k * (c0 + c1) ** 0.5)
    insz = (2 * h, 2 * w) if up else None
    ref = _ref_conv(x, wt, stride, insz, x1)
    wp = ops.split_bf16(ops.pack_weight(wt.to(DEV), torch.float32), parts)
    wup = ops.split_bf16(ops.pack_upconv2x_weight(wt.to(DEV), torch.float32), parts) if up else None
    xd, x1d = x.to(DEV), (x1.to(DEV) if x1 is not None else None)
    y = ops.conv2d(xd, wp, cout, k, stride, x1=x1d, in_size=insz, weight_up2x=wup)
    tol = 5e-5 if parts == 2 else 1e-5
    assert y.dtype == torch.float32 and relerr(y.cpu(), ref) < tol, relerr(y.cpu(), ref)
    # weight gradient against autograd in float64
    dy = torch.randn(ref.shape, generator=g).float()
    wr = wt.double().clone().requires_grad_(True)
    a = x.double().permute(0, 3, 1, 2)
    if up:
        a = F.interpolate(a, size=insz)
    if x1 is not None:
        a = torch.cat([a, x1.double().permute(0, 3, 1, 2)], 1)
    F.conv2d(a, wr, None, stride, k // 2).backward(dy.double().permute(0, 3, 1, 2))
    dw = ops.conv2d_wgrad(xd, dy.to(DEV), k, stride, x1=x1d, in_size=insz, x3=parts)
    gw = torch.empty(cout, c0 + c1, k, k, device=DEV)
    ops.unpack_wgrad(dw, gw)
    assert relerr(gw.cpu(), wr.grad) < tol, relerr(gw.cpu(), wr.grad)


# ----------------------------------------------------------------------------- BASELINE configs[0]
@pytest.mark.parametrize('precision', ['bf16x3', 'bf16x6'])
def test_config1_forward_tensor_core_parity(precision):
    """1 x 320 x 576, sparse radar input, eval: the tcgen05 engines in parity mode vs the oracle, <= 1e-3."""
    p = synth_fusionnet_state(synth.CANONICAL_FUSIONNET, 0)
    image, depth = synth.fusionnet_inputs(1, 320, 576, 0, 'sparse')
    with torch.no_grad():
        d_o, l_o = fo.fusionnet_forward(p, image, depth)
    m = make_model(synth.CANONICAL_FUSIONNET, p, precision)
    m.eval()
    with torch.no_grad():
        l = m.forward(image.to(DEV), depth.to(DEV), return_logits=True)
        d = m.forward(image.to(DEV), depth.to(DEV))
    el, ed, mae = relerr(l.cpu(), l_o), relerr(d.cpu(), d_o), float((d.cpu() - d_o).abs().mean())
    print('%s config 1: logits relerr %.2e depth relerr %.2e MAE %.2e m' % (precision, el, ed, mae))
    record('config1_%s' % precision, {'logit_relerr': el, 'depth_relerr': ed, 'depth_mae_m': mae})
    assert el < TOL and ed < TOL and mae < 1e-3


# ----------------------------------------------------------------------------- BASELINE configs[1] shape
def _oracle_train_step(p0, image, depth, gt, lidar, dtype=torch.float32):
    po = {k: (v.to(dtype) if v.is_floating_point() else v.clone()).detach().clone()
          .requires_grad_('running' not in k and v.is_floating_point()) for k, v in p0.items()}
    stats = {}
    gt_o = fo.outlier_removal(gt, 7, 1.5)
    d_o, _ = fo.fusionnet_forward(po, image.to(dtype), depth.to(dtype), training=True, new_stats=stats)
    loss_o = fo.fusionnet_loss(d_o, gt_o.to(dtype), lidar.to(dtype), 2.0, 'l1')
    loss_o.backward()
    return po, stats, d_o.detach(), float(loss_o.detach())


def _product_train_step(p0, precision, image, depth, gt, lidar):
    import net_utils
    m = make_model(synth.CANONICAL_FUSIONNET, p0, precision)
    m.train()
    gt_d = net_utils.OutlierRemoval(7, 1.5).remove_outliers(gt.to(DEV))
    d = m.forward(image.to(DEV), depth.to(DEV))
    loss, _ = m.compute_loss(image=image.to(DEV), output_depth=d, ground_truth=gt_d, lidar_map=lidar.to(DEV),
                             loss_func='l1', w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                             validity_map_loss_smoothness=torch.ones_like(gt_d), w_lidar_loss=2.0)
    loss.backward()
    torch.cuda.synchronize()
    named = dict([('encoder.' + k, v) for k, v in m.encoder.named_parameters()] +
                 [('decoder.' + k, v) for k, v in m.decoder.named_parameters()])
    grads = {k: (None if v.grad is None else v.grad.detach().cpu()) for k, v in named.items()}
    sd = {('encoder.' + k): v.detach().cpu() for k, v in m.encoder.state_dict().items()}
    sd.update({('decoder.' + k): v.detach().cpu() for k, v in m.decoder.state_dict().items()})
    return d.detach().cpu(), float(loss), grads, sd


def test_train_step_352x704_b2_tensor_core_parity():
    """One full training step at 352 x 704, batch 2 (forward, outlier removal, masked L1, backward) through the tcgen05
    engines in 'bf16x6' mode vs the oracle.
      * depth, loss, BatchNorm running statistics: <= 1e-3 against the fp32 oracle (measured ~5e-6).
      * EVERY parameter gradient: the oracle is evaluated in fp32 (what the reference computes) AND in float64 (the exact
        value of the same algorithm).  At this size the reference's own fp32 gradient deviates from the float64 one by
        4e-3 (median over the 209 tensors, max-norm relative) up to 3e-2 -- BatchNorm backward cancels large sums over
        5e5 pixels -- so "within 5e-3 of the fp32 oracle" is below the reference's own rounding noise.  The assertion is
        therefore made against the float64 gradient, distribution against distribution: median, 90th percentile and
        maximum of the per-tensor deviations each <= 1.5x the fp32 oracle's own, every tensor <= 6e-2 and cosine > 0.9995.
    The fp32 SIMT mode is held to the same bar; the bf16 fast mode's deviation on the same step is measured, bounded at
    a stated value and recorded."""
    n, h, w, seed = 2, 352, 704, 21
    p0 = synth_fusionnet_state(synth.CANONICAL_FUSIONNET, 0)
    image, depth = synth.fusionnet_inputs(n, h, w, seed, 'quasi_dense')
    gt, lidar = synth.training_targets(n, h, w, seed)
    po, stats, d_o, loss_o = _oracle_train_step(p0, image, depth, gt, lidar)
    p64, _, d_64, loss_64 = _oracle_train_step(p0, image, depth, gt, lidar, torch.float64)
    e_ref = {k: relerr(po[k].grad, p64[k].grad) for k in po if po[k].grad is not None}
    ref_sorted = sorted(e_ref.values())
    summary = {'oracle_fp32_vs_float64': {'grad_relerr_median': ref_sorted[len(ref_sorted) // 2], 'grad_relerr_max': ref_sorted[-1],
                                          'depth_relerr': relerr(d_o, d_64), 'loss_relerr': abs(loss_o - loss_64) / abs(loss_64)}}
    print('oracle fp32 vs float64', {k: '%.2e' % v for k, v in summary['oracle_fp32_vs_float64'].items()})
    res = {prec: _product_train_step(p0, prec, image, depth, gt, lidar) for prec in ('fp32', 'bf16x6', 'bf16')}
    all_errs = {}
    for prec, (d, loss, grads, sd) in res.items():
        errs = {}
        for k in grads:
            assert (grads[k] is None) == (po[k].grad is None), (prec, k)
            if grads[k] is not None:
                errs[k] = relerr(grads[k], p64[k].grad)
        srt = sorted(errs.values())
        cos = sorted(float(F.cosine_similarity(grads[k].flatten().double(), p64[k].grad.flatten(), dim=0))
                     for k in errs if grads[k].numel() >= 1024)
        summary[prec] = {
            'depth_relerr': relerr(d, d_o), 'depth_mae_m': float((d - d_o).abs().mean()),
            'depth_max_m': float((d - d_o).abs().max()), 'loss_relerr': abs(loss - loss_o) / abs(loss_o),
            'grad_relerr_vs_float64_median': srt[len(srt) // 2], 'grad_relerr_vs_float64_max': srt[-1],
            'grad_cosine_min': cos[0], 'grad_cosine_median': cos[len(cos) // 2],
            'bn_running_relerr_max': max(relerr(sd[k], v) for k, v in stats.items()), 'n_grad_tensors': len(errs)}
        all_errs[prec] = errs
        print(prec, {k: ('%.2e' % v if isinstance(v, float) else v) for k, v in summary[prec].items()})
    record('train_step_352x704_b2', summary)
    ref_med = summary['oracle_fp32_vs_float64']['grad_relerr_median']
    for prec in ('bf16x6', 'fp32'):
        sp, ep = summary[prec], all_errs[prec]
        assert sp['depth_relerr'] < TOL and sp['depth_mae_m'] < 1e-3 and sp['loss_relerr'] < TOL, (prec, sp)
        assert sp['bn_running_relerr_max'] < TOL, (prec, sp)
        # the two noise realisations are independent, so they are compared as distributions over the 209 tensors
        # (median, 90th percentile, maximum: each <= 1.5x the fp32 oracle's own), with a per-tensor cap and a direction check
        ours, theirs = sorted(ep.values()), ref_sorted
        for q in (0.5, 0.9, 1.0):
            i = min(int(q * len(ours)), len(ours) - 1)
            assert ours[i] < max(1.5 * theirs[i], 2e-3), (prec, q, ours[i], theirs[i])
        assert ours[-1] < 6e-2 and sp['grad_cosine_min'] > 0.9995, (prec, sp)
    # fast mode: recorded above; stated bounds (bf16 rounding of activations and weights, ~35 layers deep)
    sb = summary['bf16']
    assert sb['depth_mae_m'] < 0.06 and sb['loss_relerr'] < 1e-2 and sb['grad_cosine_median'] > 0.9 and sb['grad_cosine_min'] > 0.8


# ----------------------------------------------------------------------------- BASELINE configs[2] shape
def _radarnet_case(k, seed, h=352, w=704, ph=352, pw=288):
    import radarnet_model
    cfg = dict(synth.CANONICAL_RADARNET, input_patch_size_image=(ph, pw))
    m = radarnet_model.RadarNetModel(device=DEV, **cfg)
    p = {}
    for kk, v in m.encoder.state_dict().items():
        p['encoder.' + kk] = v
    for kk, v in m.decoder.state_dict().items():
        p['decoder.' + kk] = v
    synth.fill_state_dict_(p, seed)                 # in place on the CUDA parameters (same CPU generator values)
    m.eval()
    pad = pw // 2
    gen = torch.Generator().manual_seed(seed)
    image = torch.rand(1, 3, h, w + 2 * pad, generator=gen)
    pt = synth.radar_points(k, h, w, seed)
    pt[:, 0] += pad
    boxes = [torch.stack([pt[:, 0] - pad, torch.zeros(k), pt[:, 0] + pad, torch.full((k,), float(h))], 1)]
    return m, {kk: v.detach().cpu() for kk, v in p.items()}, image, pt, boxes


def test_radarnet_352x704_k64_tensor_core_parity():
    """RadarNet stage-1 column at the BASELINE configs[2] shape (352 x 704 image, 64 radar points, 352 x 288 patches):
    logits of all 64 crops through the tcgen05 engines in 'bf16x3' mode vs the oracle <= 1e-3 (relative to the max
    logit); the bf16 fast mode (what tools/bench_radarnet.py and bench.py --mode radarnet time) is measured, bounded
    and recorded, and its S2 scatter output is compared with the scatter of the oracle's crops."""
    import scatter_oracle as so
    k = 64
    m, p_cpu, image, pt, boxes = _radarnet_case(k, 4)
    with torch.no_grad():
        ref = ro.radarnet_forward(p_cpu, image, pt, boxes, (352, 288), roi_pool=ro.roi_pool_vectorised)
    out = {}
    for prec in ('bf16x3', 'bf16'):
        m.set_precision(prec)
        with torch.no_grad():
            out[prec] = m.forward(image.to(DEV), pt.to(DEV), boxes, return_logits=True).cpu()
    e3 = relerr(out['bf16x3'], ref)
    eb = relerr(out['bf16'], ref)
    sig = lambda t: torch.sigmoid(t)
    resp_mae = float((sig(out['bf16']) - sig(ref)).abs().mean())
    resp_max = float((sig(out['bf16']) - sig(ref)).abs().max())
    # S2 on the fast path's crops vs S2 on the oracle's crops: share of pixels whose arg-max point differs
    d_b, r_b = so.s2_scatter(sig(out['bf16']).numpy(), pt.numpy(), 704, (352, 288), compat=False)
    d_r, r_r = so.s2_scatter(sig(ref).numpy(), pt.numpy(), 704, (352, 288), compat=False)
    occ_diff = float(((r_b != 0) != (r_r != 0)).mean())
    depth_diff = float((d_b != d_r).mean())
    print('radarnet 352x704 k64: bf16x3 logit relerr %.2e; bf16 logit relerr %.2e, response MAE %.2e max %.2e, '
          'S2 occupancy differs on %.3f %% of pixels, depth on %.3f %%' % (e3, eb, resp_mae, resp_max, 100 * occ_diff, 100 * depth_diff))
    record('radarnet_352x704_k64', {'bf16x3_logit_relerr': e3, 'bf16_logit_relerr': eb, 'bf16_response_mae': resp_mae,
                                    'bf16_response_max': resp_max, 'bf16_s2_occupancy_diff_frac': occ_diff,
                                    'bf16_s2_depth_diff_frac': depth_diff})
    assert e3 < TOL
    assert eb < 5e-2 and resp_mae < 5e-3
