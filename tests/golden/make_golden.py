"""
Generates tests/golden/*.npz by executing the UNMODIFIED reference (imported from
/root/reference/src, CPU fp32) on deterministic synthetic inputs, and checks the CPU
oracle against it on the spot.  Run in the build container only (the GPU box has no
/root/reference):

    python tests/golden/make_golden.py

Weights are never shipped for the canonical (14.4 M parameter) config: both this script
and the tests regenerate them with rcfd.synth.fill_state_dict_(state_dict, seed), which
walks the state_dict in its own (reference-defined) order.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get('RCFD_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))

# the only missing import of the reference: log_utils.py:17 -> matplotlib
for name in ('matplotlib', 'matplotlib.pyplot'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


def _import_reference():
    # The reference's module names (networks, net_utils, ...) collide with the product's
    # flat modules, so import them with the reference's src dir FIRST on sys.path and
    # keep handles.
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(REF, 'src'))
    import fusionnet_model as ref_fm
    import radarnet_model as ref_rm
    import radarnet_main as ref_rmain
    import net_utils as ref_nu
    mods = dict(fusionnet_model=ref_fm, radarnet_model=ref_rm, radarnet_main=ref_rmain, net_utils=ref_nu)
    sys.path[:] = saved
    for m in ('fusionnet_model', 'radarnet_model', 'radarnet_main', 'net_utils', 'networks',
              'fusionnet_losses', 'log_utils', 'radarnet_transforms', 'datasets', 'data_utils',
              'eval_utils'):
        sys.modules.pop(m, None)
    return mods


REFM = _import_reference()
from rcfd import synth                      # noqa: E402
import fusionnet_oracle as fo               # noqa: E402
import radarnet_oracle as ro                # noqa: E402
import scatter_oracle as so                 # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def flat_state(model):
    p = {}
    for k, v in model.encoder.state_dict().items():
        p['encoder.' + k] = v
    for k, v in model.decoder.state_dict().items():
        p['decoder.' + k] = v
    return p


def build_fusionnet(cfg, seed):
    model = REFM['fusionnet_model'].FusionNetModel(device=torch.device('cpu'), **cfg)
    p = flat_state(model)
    synth.fill_state_dict_(p, seed)      # in place on the module's own storage
    return model, p


def fusionnet_case(name, cfg, n, h, w, seed, variant, save_weights, train=True):
    model, p = build_fusionnet(cfg, seed)
    image, depth = synth.fusionnet_inputs(n, h, w, seed, variant)
    out = {}
    # ---- eval mode
    model.eval()
    with torch.no_grad():
        latent, skips = model.encoder(image=image, depth=depth)
        logits = model.decoder(x=latent, skips=skips, shape=image.shape[-2:])[-1]
        d_ref = model.forward(image, depth)
        d_or, l_or = fo.fusionnet_forward(p, image, depth, n_levels=len(cfg['n_filters_encoder_image']))
    assert relerr(l_or, logits) < 1e-5 and relerr(d_or, d_ref) < 1e-5, (relerr(l_or, logits), relerr(d_or, d_ref), float(logits.abs().max()))
    out.update(eval_logits=logits.numpy(), eval_depth=d_ref.numpy(), eval_latent=latent.numpy(),
               eval_skip1=skips[0].numpy()[:, :4], eval_skip3=skips[2].numpy()[:, :4])
    if save_weights:
        for k, v in p.items():
            out['w::' + k] = v.numpy().copy()
    out['meta'] = np.array([n, h, w, seed])
    out['variant'] = np.array(variant)
    if not train:
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print(name, 'ok (eval only) depth range', float(d_ref.min()), float(d_ref.max()),
              'logit range', float(logits.min()), float(logits.max()))
        return
    # ---- train mode: batch statistics, loss, grads, running-stat update
    gt, lidar = synth.training_targets(n, h, w, seed)
    p_before = {k: v.clone() for k, v in p.items()}
    model.train()
    for q in model.parameters():
        q.grad = None
    d_tr = model.forward(image, depth)
    gt_clean = REFM['net_utils'].OutlierRemoval(7, 1.5).remove_outliers(gt)
    loss, _ = model.compute_loss(image=image, output_depth=d_tr, ground_truth=gt_clean, lidar_map=lidar,
                                 loss_func='l1', w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                                 validity_map_loss_smoothness=torch.ones_like(gt), w_lidar_loss=2.0)
    loss.backward()
    p_after = flat_state(model)
    # oracle on the same
    po = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in p_before.items()}
    new_stats = {}
    d_or, _ = fo.fusionnet_forward(po, image, depth, n_levels=len(cfg['n_filters_encoder_image']),
                                   training=True, new_stats=new_stats)
    gt_or = fo.outlier_removal(gt, 7, 1.5)
    assert torch.equal(gt_or, gt_clean)
    loss_or = fo.fusionnet_loss(d_or, gt_or, lidar, 2.0, 'l1')
    loss_or.backward()
    assert relerr(d_or.detach(), d_tr.detach()) < 1e-5
    assert abs(float(loss_or.detach()) - float(loss.detach())) < 1e-5 * abs(float(loss.detach()))
    names = [k for k in p_before if po[k].requires_grad]
    ref_named = dict(list(('encoder.' + k, v) for k, v in model.encoder.named_parameters()) +
                     list(('decoder.' + k, v) for k, v in model.decoder.named_parameters()))
    gsum, gabs, ghead, gnone = [], [], [], []
    for k in names:
        g_ref = ref_named[k].grad
        g_or = po[k].grad
        assert (g_ref is None) == (g_or is None), k
        if g_ref is None:
            gnone.append(1); gsum.append(0.0); gabs.append(0.0); ghead.append(np.zeros(8, np.float32))
            continue
        assert relerr(g_or, g_ref) < 1e-4, (k, relerr(g_or, g_ref))
        gnone.append(0)
        gsum.append(float(g_ref.double().sum())); gabs.append(float(g_ref.double().abs().sum()))
        ghead.append(g_ref.flatten()[:8].numpy().astype(np.float32).copy() if g_ref.numel() >= 8
                     else np.resize(g_ref.flatten().numpy(), 8).astype(np.float32))
    for k, v in new_stats.items():
        assert relerr(v, p_after[k]) < 1e-5, k
    rm_key = 'decoder.deconv0.conv.batch_norm.running_mean'
    out.update(train_depth=d_tr.detach().numpy(), train_loss=np.float64(float(loss)),
               grad_names=np.array(names), grad_sum=np.array(gsum), grad_abs=np.array(gabs),
               grad_head=np.stack(ghead), grad_none=np.array(gnone),
               train_running_mean_deconv0=p_after[rm_key].numpy(),
               train_running_var_deconv0=p_after[rm_key.replace('mean', 'var')].numpy(),
               gt_clean=gt_clean.numpy())
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'ok  loss', float(loss), 'depth range', float(d_ref.min()), float(d_ref.max()),
          'logit range', float(logits.min()), float(logits.max()))


def radarnet_case(name, cfg, n, h, w, k, seed):
    model = REFM['radarnet_model'].RadarNetModel(device=torch.device('cpu'), **cfg)
    p = flat_state(model)
    synth.fill_state_dict_(p, seed)
    model.eval()
    ph, pw = cfg['input_patch_size_image']
    pad = pw // 2
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(n, 3, h, w + 2 * pad, generator=g)       # already edge-padded size
    pts, boxes = [], []
    for b in range(n):
        pt = synth.radar_points(k, h, w, seed + b)
        pt[:, 0] += pad                                           # radarnet_main.py:980-983
        bx = torch.stack([pt[:, 0] - pad, torch.zeros(k), pt[:, 0] + pad, torch.full((k,), float(h))], 1)
        pts.append(pt); boxes.append(bx)
    points = torch.cat(pts, 0)
    with torch.no_grad():
        logits = model.forward(image, points, boxes, return_logits=True)
        l_or = ro.radarnet_forward(p, image, points, boxes, (ph, pw),
                                   n_filters_image=cfg['n_filters_encoder_image'],
                                   n_neuron_latent=cfg['n_neurons_encoder_depth'][-1])
    assert relerr(l_or, logits) < 1e-5, relerr(l_or, logits)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), logits=logits.numpy(),
                        meta=np.array([n, h, w, k, seed, ph, pw]))
    print(name, 'ok  logits range', float(logits.min()), float(logits.max()))


def roi_pool_case():
    import torchvision
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 3, 22, 62, generator=g)
    boxes = [torch.tensor([[3.2, 0.0, 291.2, 352.0], [100.5, 0.0, 388.5, 352.0]]),
             torch.tensor([[650.9, 0.0, 938.9, 352.0]])]
    for scale, osz in ((1 / 16.0, (22, 18)), (1 / 32.0, (11, 9))):
        f = feat if scale == 1 / 16.0 else feat[:, :, :11, :31].contiguous()
        ref = torchvision.ops.roi_pool(f, boxes, spatial_scale=scale, output_size=osz)
        mine = ro.roi_pool(f, boxes, scale, osz)
        assert torch.equal(ref, mine), 'roi_pool oracle mismatch'
    print('roi_pool oracle == torchvision')


class _StubModel(object):
    """Feeds pre-baked crops through the real radarnet_main.forward (S2 golden)."""
    def __init__(self, crops, patch):
        self.crops = crops
        self.input_patch_size_image = patch

    def forward(self, image, point, bounding_boxes, return_logits):
        return self.crops


def s2_case(name, h, w, patch, k, seed, zs=None):
    g = torch.Generator().manual_seed(seed)
    ph, pw = patch
    pad = pw // 2
    crops = torch.rand(k, 1, ph, pw, generator=g)
    crops[:, :, : ph // 3] *= 0.4                      # plenty of sub-threshold pixels
    pts = synth.radar_points(max(k, 8), h, w, seed)[:k].clone()
    if zs is not None:
        pts[:, 2] = torch.tensor(zs)
    pts[:, 0] += pad
    image = torch.zeros(1, 3, h, w)
    depth, resp = REFM['radarnet_main'].forward(_StubModel(crops, patch), image, pts.clone(), None,
                                                device=torch.device('cpu'))
    d_or, r_or = so.s2_scatter(crops.numpy(), pts.numpy(), w, patch, compat=True)
    assert np.array_equal(d_or, depth.numpy()) and np.array_equal(r_or, resp.numpy())
    np.savez_compressed(os.path.join(OUT, name + '.npz'), crops=crops.numpy(), points=pts.numpy(),
                        depth=depth.numpy(), response=resp.numpy(), meta=np.array([h, w, ph, pw, k, seed]))
    print(name, 'ok  nonzero', int((depth != 0).sum()), 'dtype', depth.dtype)


def png16_case(name, h, w, seed):
    """The PNG round trip between the two stages, executed by the reference's own codec (src/data_utils.py:238-335)."""
    import tempfile
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(REF, 'src'))
    import data_utils as ref_du
    sys.path[:] = saved
    sys.modules.pop('data_utils', None)
    rng = np.random.default_rng(seed)
    depth_i64 = rng.integers(0, 100, (h, w)).astype(np.int64) * (rng.random((h, w)) < 0.3)
    depth_f32 = (rng.random((h, w)) * 100).astype(np.float32) * (rng.random((h, w)) < 0.3)
    response = rng.random((h, w)).astype(np.float32) * (rng.random((h, w)) < 0.3)
    tmp = tempfile.mkdtemp()
    out = dict(depth_i64=depth_i64, depth_f32=depth_f32, response=response)
    for tag, dep in (('i64', depth_i64), ('f32', depth_f32)):
        ref_du.save_depth(dep, os.path.join(tmp, 'd.png'))
        ref_du.save_response(response, os.path.join(tmp, 'r.png'))
        out['loaded_depth_' + tag] = ref_du.load_depth(os.path.join(tmp, 'd.png'))
        out['loaded_response'] = ref_du.load_response(os.path.join(tmp, 'r.png'))
        d_o, r_o = so.png16_roundtrip(dep, response)
        assert np.array_equal(d_o, out['loaded_depth_' + tag]) and np.array_equal(r_o, out['loaded_response'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'ok (oracle == reference codec, bit exact)')


def transforms_case(name, seed):
    """The on-device augmentation of the reference (src/fusionnet_transforms.py:46-178) on seeded inputs, executed by
    the reference itself; the oracle must replay its random draws and reproduce every value bit for bit."""
    import transforms_oracle as tro
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(REF, 'src'))
    import fusionnet_transforms as ref_tr
    sys.path[:] = saved
    sys.modules.pop('fusionnet_transforms', None)
    cfg = dict(random_brightness=[0.8, 1.2], random_contrast=[0.8, 1.2], random_saturation=[0.8, 1.2],
               random_flip_type=['horizontal', 'vertical'])
    out = {'meta': np.array([seed])}
    for tag, scale, rng in (('u8_01', 255.0, [0, 1]), ('f_pm1', 1.0, [-1, 1])):
        g = torch.Generator().manual_seed(seed)
        img = torch.rand(5, 3, 18, 26, generator=g) * scale
        if scale > 1.0:
            img = img.round()
        maps = [torch.rand(5, 1, 18, 26, generator=g) * 50, torch.rand(5, 2, 18, 26, generator=g)]
        t = ref_tr.Transforms(normalized_image_range=rng, **cfg)
        torch.manual_seed(seed + 1)
        (ri,), rr = t.transform([img.clone()], [m.clone() for m in maps], random_transform_probability=0.9)
        torch.manual_seed(seed + 1)
        d = tro.draw_decisions(5, 0.9, cfg['random_brightness'], cfg['random_contrast'], cfg['random_saturation'],
                               cfg['random_flip_type'])
        oi, orr = tro.apply(img, maps, d, rng)
        assert torch.equal(oi, ri) and all(torch.equal(a, b) for a, b in zip(orr, rr))
        out[tag + '_image'] = ri.numpy()
        out[tag + '_map0'] = rr[0].numpy()
        out[tag + '_map1'] = rr[1].numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'ok (oracle == reference transforms, bit exact)')


def losses_case(name, seed):
    """fusionnet_losses.* (src/fusionnet_losses.py) and the non-canonical branches of FusionNetModel.compute_loss
    (src/fusionnet_model.py:172-302: l2 / smoothl1, first-order and Sobel smoothness, with and without the lidar term)
    evaluated by the reference on seeded inputs."""
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(REF, 'src'))
    import fusionnet_losses as ref_l
    sys.path[:] = saved
    sys.modules.pop('fusionnet_losses', None)
    g = torch.Generator().manual_seed(seed)
    n, h, w = 2, 24, 40
    image = torch.rand(n, 3, h, w, generator=g)
    pred = torch.rand(n, 1, h, w, generator=g) * 50 + 1
    tgt = torch.rand(n, 1, h, w, generator=g) * 50 + 1
    weights = (torch.rand(n, 1, h, w, generator=g) > 0.3).float()
    gt = tgt * (torch.rand(n, 1, h, w, generator=g) < 0.4)
    lidar = (torch.rand(n, 1, h, w, generator=g) * 50 + 1) * (torch.rand(n, 1, h, w, generator=g) < 0.05)
    out = {'meta': np.array([seed, n, h, w])}
    out['l1'] = ref_l.l1_loss(pred, tgt).numpy()
    out['l2'] = ref_l.l2_loss(pred, tgt).numpy()
    out['smoothl1'] = ref_l.smooth_l1_loss(pred, tgt).numpy()
    out['smooth'] = ref_l.smoothness_loss_func(pred, image).numpy()
    out['sobel7'] = ref_l.sobel_smoothness_loss_func(pred, image, weights, [1, 1, 7, 7]).numpy()
    out['sobel3'] = ref_l.sobel_smoothness_loss_func(pred, image, weights, [1, 1, 3, 3]).numpy()
    gx, gy = ref_l.sobel_filter([1, 1, 7, 7])
    out['sobel_gx'], out['sobel_gy'] = gx.numpy(), gy.numpy()
    dy, dx = ref_l.gradient_yx(pred)
    out['grad_dy'], out['grad_dx'] = dy.numpy(), dx.numpy()
    model, _ = build_fusionnet(synth.SMALL_FUSIONNET, 1)
    combos = [('l1', 0.0, -1, 2.0), ('l2', 0.0, -1, 2.0), ('smoothl1', 0.0, -1, 0.0), ('l1', 0.5, -1, 2.0),
              ('l1', 0.5, 7, 2.0), ('l2', 0.25, 3, 0.0)]
    vals = []
    for lf, ws, ks, wl in combos:
        loss, _ = model.compute_loss(image=image, output_depth=pred, ground_truth=gt, lidar_map=lidar, loss_func=lf,
                                     w_smoothness=ws, loss_smoothness_kernel_size=ks,
                                     validity_map_loss_smoothness=weights, w_lidar_loss=wl)
        vals.append(float(loss))
    out['compute_loss'] = np.array(vals)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'ok', vals)


def _reference_statements(path, func_name, first_line, last_line):
    """The statements of ``func_name`` in the reference file ``path`` that lie entirely inside [first_line, last_line]
    (outermost ones only), compiled as a module: lets a slice of a function that cannot be called as a whole (nuScenes
    objects) be EXECUTED as it stands in the reference, on our own inputs."""
    import ast
    tree = ast.parse(open(path).read())
    func = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == func_name)
    picked = []

    def visit(stmts):
        for st in stmts:
            if st.lineno >= first_line and st.end_lineno <= last_line:
                picked.append(st)
            elif st.lineno <= last_line and st.end_lineno >= first_line:
                for field in ('body', 'orelse', 'finalbody'):
                    visit(getattr(st, field, []) or [])
    visit(func.body)
    assert picked, (func_name, first_line, last_line)
    return compile(ast.Module(body=picked, type_ignores=[]), path, 'exec')


def s1_case(name, h, w, seed):
    """Scatter S1 executed by the reference's own code: `points_to_depth_map` is called as a function; the main-sweep
    plot, the z-buffer merge of a later sweep and the final nonzero -> point list are the reference's statements
    (setup/setup_dataset_nuscenes_with_denseGT.py:641-659, :699-713, :771-782) run on seeded points."""
    import ast
    path = os.path.join(REF, 'setup', 'setup_dataset_nuscenes_with_denseGT.py')
    tree = ast.parse(open(path).read())
    fdef = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == 'points_to_depth_map')
    ns = {'np': np}
    exec(compile(ast.Module(body=[fdef], type_ignores=[]), path, 'exec'), ns)
    rng = np.random.default_rng(seed)

    def sweep(n):
        xy = np.stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)])
        xy[:, :4] = np.array([[10.5, 11.5, 30.5, 31.5], [20.5, 21.5, 7.5, 8.5]])[:, :min(4, n)]   # half-even ties
        xy[:, 5] = xy[:, 4]                                                                   # duplicate pixel
        return xy, rng.uniform(1, 80, n)
    sweeps = [sweep(40), sweep(30), sweep(30)]
    sweeps[1][0][:, :10] = sweeps[0][0][:, :10]             # collisions with the main sweep: closer / farther points
    sweeps[1][1][:5] = sweeps[0][1][:5] * 0.5
    sweeps[1][1][5:10] = sweeps[0][1][5:10] * 2.0
    image = np.zeros((h, w, 3))
    plot = ns['points_to_depth_map'](sweeps[0][0], sweeps[0][1], image)
    assert np.array_equal(plot, so.s1_points_to_depth_map(sweeps[0][0], sweeps[0][1], h, w))
    # merge_radar_point_clouds: main sweep
    env = {'np': np, 'main_image': image, 'main_points_radar': sweeps[0][0], 'main_depth_radar': sweeps[0][1]}
    exec(_reference_statements(path, 'merge_radar_point_clouds', 641, 659), env)
    merge_code = _reference_statements(path, 'merge_radar_point_clouds', 699, 713)
    for xy, z in sweeps[1:]:
        env['next_points_radar_main'], env['next_depth_radar_main'] = xy, z
        exec(merge_code, env)
    exec(_reference_statements(path, 'merge_radar_point_clouds', 771, 780), env)
    img_o, pts_o, dep_o = so.s1_merge(sweeps, h, w)
    assert np.array_equal(env['main_radar_image'], img_o)
    assert np.array_equal(env['return_points_radar'], pts_o) and np.array_equal(env['return_depth_radar'], dep_o)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), meta=np.array([h, w, seed]),
                        **{'xy%d' % i: sw[0] for i, sw in enumerate(sweeps)}, **{'z%d' % i: sw[1] for i, sw in enumerate(sweeps)},
                        plot=plot, merged=env['main_radar_image'], points=env['return_points_radar'],
                        depth=env['return_depth_radar'])
    print(name, 'ok (oracle == reference statements, bit exact); occupied pixels', int((env['main_radar_image'] > 0).sum()))


def radarnet_loss_case(name, seed):
    """RadarNetModel.compute_loss (weighted BCE over valid pixels, src/radarnet_model.py:126-167) evaluated by the reference."""
    model = REFM['radarnet_model'].RadarNetModel(device=torch.device('cpu'),
                                                 **dict(synth.CANONICAL_RADARNET, input_patch_size_image=(64, 64)))
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(3, 1, 64, 64, generator=g) * 3
    gt = (torch.rand(3, 1, 64, 64, generator=g) < 0.2).float()
    valid = (torch.rand(3, 1, 64, 64, generator=g) < 0.7).float()
    vals = [float(model.compute_loss(logits, gt, valid, w_positive_class=w)[0]) for w in (1.0, 2.0, 5.5)]
    np.savez_compressed(os.path.join(OUT, name + '.npz'), meta=np.array([seed]), loss=np.array(vals))
    print(name, 'ok', vals)


class _DummyDepthModel(object):
    """Stands in for FusionNetModel in the validate() fixture: a deterministic function of its inputs."""

    def forward(self, image, input_depth):
        return 1.0 + 40.0 * image.mean(dim=1, keepdim=True) + 0.5 * input_depth[:, 0:1]


def validate_case(name, seed):
    """The reference's validation loop, metrics and best-result rule (src/fusionnet_main.py:476-606,
    src/eval_utils.py) on seeded samples with a dummy model; two passes so that the best-result update is exercised."""
    import tempfile
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(REF, 'src'))
    import fusionnet_main as ref_main
    import fusionnet_transforms as ref_tr
    sys.path[:] = saved
    for m in ('fusionnet_main', 'fusionnet_transforms', 'fusionnet_model', 'networks', 'net_utils', 'fusionnet_losses',
              'log_utils', 'datasets', 'data_utils', 'eval_utils'):
        sys.modules.pop(m, None)
    g = torch.Generator().manual_seed(seed)
    loader = []
    for _ in range(3):
        image = (torch.rand(1, 3, 20, 32, generator=g) * 255).round()
        depth = torch.rand(1, 1, 20, 32, generator=g) * 60 * (torch.rand(1, 1, 20, 32, generator=g) < 0.2)
        response = torch.rand(1, 1, 20, 32, generator=g)
        gt = torch.rand(1, 1, 20, 32, generator=g) * 90 * (torch.rand(1, 1, 20, 32, generator=g) < 0.5)
        loader.append((image, depth, response, gt))
    tr = ref_tr.Transforms(normalized_image_range=[0, 1])
    log_path = os.path.join(tempfile.mkdtemp(), 'log.txt')
    best = {'step': -1, 'mae': np.infty if hasattr(np, 'infty') else np.inf, 'rmse': np.inf, 'imae': np.inf, 'irmse': np.inf}
    best = ref_main.validate(_DummyDepthModel(), loader, tr, step=10, best_results=best, min_evaluate_depth=1.0,
                             max_evaluate_depth=80.0, device=torch.device('cpu'), summary_writer=None, log_path=log_path)
    first = [best['step'], best['mae'], best['rmse'], best['imae'], best['irmse']]
    worse = dict(best)
    worse.update(mae=best['mae'] - 1.0, rmse=best['rmse'] - 1.0)          # 2 of 4 not better -> no update
    second = ref_main.validate(_DummyDepthModel(), loader, tr, step=20, best_results=dict(worse), min_evaluate_depth=1.0,
                               max_evaluate_depth=80.0, device=torch.device('cpu'), summary_writer=None, log_path=log_path)
    out = dict(meta=np.array([seed]), first=np.array(first, dtype=np.float64),
               second=np.array([second['step'], second['mae'], second['rmse'], second['imae'], second['irmse']], dtype=np.float64),
               log=np.array(open(log_path).read()))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'ok', first, second['step'])


CROP_TYPES = (['none'], ['center'], ['left', 'top'], ['right', 'bottom'], ['horizontal'], ['horizontal', 'anchored'],
              ['horizontal', 'vertical'], ['horizontal', 'vertical', 'anchored'], ['bottom', 'horizontal'])


def crop_case(name):
    """Crop origins chosen by the reference's own random_crop (src/datasets.py:19-109) for every crop type under seeded
    numpy draws; rcfd.data.crop_origin must make the same decisions (checked here and in tests/test_host_logic.py)."""
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(REF, 'src'))
    import datasets as ref_ds
    sys.path[:] = saved
    for m in ('datasets', 'data_utils'):
        sys.modules.pop(m, None)
    from rcfd import data as prod_data
    oh, ow, nh, nw = 90, 160, 35, 70
    index = np.arange(oh * ow, dtype=np.float32).reshape(1, oh, ow)
    origins = np.zeros((len(CROP_TYPES), 8, 2), dtype=np.int32)
    for ci, crop_type in enumerate(CROP_TYPES):
        for seed in range(8):
            np.random.seed(100 + seed)
            [out] = ref_ds.random_crop([index], (nh, nw), crop_type)
            y0, x0 = divmod(int(out[0, 0, 0]), ow)
            origins[ci, seed] = (y0, x0)
            got = prod_data.crop_origin(oh, ow, nh, nw, crop_type, rng=np.random.RandomState(100 + seed))
            assert got == (y0, x0), (crop_type, seed, got, (y0, x0))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), origins=origins, meta=np.array([oh, ow, nh, nw]))
    print(name, 'ok (crop_origin == reference random_crop for', len(CROP_TYPES), 'crop types x 8 seeds)')


TINY_FUSIONNET = dict(synth.CANONICAL_FUSIONNET, n_filters_encoder_image=[8, 8, 16, 16, 16, 16],
                      n_filters_encoder_depth=[8, 8, 8, 8, 8, 8], n_filters_decoder=[16, 16, 16, 8, 8, 8])


def checkpoint_case(name, seed):
    """Checkpoint interop in both directions.  (1) The unmodified reference -- wrapped in
    torch.nn.DataParallel like src/fusionnet_main.py:198 does before it saves -- takes one Adam step and writes
    tests/golden/<name>.pth with its own save_model; the product model + rcfd.optim.FusedAdam restore it.  (2) A
    checkpoint written by the product's save_model with FusedAdam's state loads into the reference's restore_model
    (strict DataParallel keys) and its torch.optim.Adam, which can then step."""
    import tempfile
    import fusionnet_model as prod_fm
    from rcfd import optim as prod_optim
    torch.manual_seed(seed)
    ref = REFM['fusionnet_model'].FusionNetModel(device=torch.device('cpu'), **TINY_FUSIONNET)
    synth.fill_state_dict_(flat_state(ref), seed)
    ref.train()
    ref.data_parallel()
    opt = torch.optim.Adam([{'params': ref.parameters(), 'weight_decay': 0.0}], lr=1e-3)
    image, depth = synth.fusionnet_inputs(2, 64, 128, seed, 'quasi_dense')
    gt, lidar = synth.training_targets(2, 64, 128, seed)
    out = ref.forward(image, depth)
    loss, _ = ref.compute_loss(image=image, output_depth=out, ground_truth=gt, lidar_map=lidar, loss_func='l1',
                               w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                               validity_map_loss_smoothness=torch.ones_like(gt), w_lidar_loss=2.0)
    opt.zero_grad()
    loss.backward()
    opt.step()
    path = os.path.join(OUT, name + '.pth')
    ref.save_model(path, 7, opt)
    ck = torch.load(path, weights_only=False)
    assert all(k.startswith('module.') for k in ck['encoder_state_dict'])
    # (1) reference file -> product
    prod = prod_fm.FusionNetModel(device=torch.device('cpu'), **TINY_FUSIONNET)
    popt = prod_optim.FusedAdam([{'params': prod.parameters(), 'weight_decay': 0.0}], lr=5e-4)
    step, popt = prod.restore_model(path, optimizer=popt)
    assert step == 7 and popt.step_count == 1 and popt.param_groups[0]['lr'] == 1e-3
    for a, b in zip(prod.parameters(), ref.parameters()):
        assert torch.equal(a, b)
    sd_ref = opt.state_dict()
    sd_prod = popt.state_dict()
    assert set(sd_prod['state']) == set(sd_ref['state']), (len(sd_prod['state']), len(sd_ref['state']))
    for i, st in sd_ref['state'].items():
        assert torch.equal(sd_prod['state'][i]['exp_avg'], st['exp_avg']) and float(sd_prod['state'][i]['step']) == float(st['step'])
    assert set(sd_ref['param_groups'][0]) <= set(sd_prod['param_groups'][0]), set(sd_ref['param_groups'][0]) - set(sd_prod['param_groups'][0])
    # (2) product file -> reference
    with tempfile.TemporaryDirectory() as tmp:
        p2 = os.path.join(tmp, 'model-9.pth')
        prod.save_model(p2, 9, popt)
        ref2 = REFM['fusionnet_model'].FusionNetModel(device=torch.device('cpu'), **TINY_FUSIONNET)
        ref2.data_parallel()
        opt2 = torch.optim.Adam([{'params': ref2.parameters(), 'weight_decay': 0.0}], lr=1e-3)
        step2, opt2 = ref2.restore_model(p2, optimizer=opt2)
        assert step2 == 9
        for a, b in zip(ref2.parameters(), ref.parameters()):
            assert torch.equal(a, b)
        ref2.train()
        out2 = ref2.forward(image, depth)
        loss2, _ = ref2.compute_loss(image=image, output_depth=out2, ground_truth=gt, lidar_map=lidar, loss_func='l1',
                                     w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                                     validity_map_loss_smoothness=torch.ones_like(gt), w_lidar_loss=2.0)
        opt2.zero_grad()
        loss2.backward()
        opt2.step()                       # the restored Adam state is usable
        assert int(float(opt2.state_dict()['state'][0]['step'])) == 2
    print(name, 'ok: reference checkpoint', os.path.getsize(path), 'bytes, both directions load')


def multires_case(name, n_resolution, n, h, w, seed):
    """Multi-resolution decoder (src/networks.py:1595-1642): every output of forward(return_multiscale=True) in train mode,
    a loss over ALL scales (so that every output{k} conv and the bilinear / concat path carry gradient) and the
    gradient fingerprints, written by the reference; the oracle is checked against it on the spot."""
    cfg = dict(synth.SMALL_FUSIONNET, n_resolution_decoder=n_resolution)
    model, p = build_fusionnet(cfg, seed)
    image, depth = synth.fusionnet_inputs(n, h, w, seed, 'quasi_dense')
    p_before = {k: v.clone() for k, v in p.items()}
    model.train()
    outs = model.forward(image, depth, return_multiscale=True)
    assert len(outs) == n_resolution
    weights = [float(i + 1) for i in range(n_resolution)]
    loss = sum(wi * o.mean() for wi, o in zip(weights, outs))
    loss.backward()
    po = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in p_before.items()}
    d_or, _ = fo.fusionnet_forward(po, image, depth, training=True, return_multiscale=True)
    loss_or = sum(wi * o.mean() for wi, o in zip(weights, d_or))
    loss_or.backward()
    for a, b in zip(d_or, outs):
        assert relerr(a.detach(), b.detach()) < 1e-5
    ref_named = dict(list(('encoder.' + k, v) for k, v in model.encoder.named_parameters()) +
                     list(('decoder.' + k, v) for k, v in model.decoder.named_parameters()))
    names, gsum, gabs, gnone = [], [], [], []
    for k, v in ref_named.items():
        g_or = po[k].grad
        assert (v.grad is None) == (g_or is None), k
        names.append(k)
        if v.grad is None:
            gnone.append(1); gsum.append(0.0); gabs.append(0.0)
            continue
        assert relerr(g_or, v.grad) < 1e-4, (k, relerr(g_or, v.grad))
        gnone.append(0); gsum.append(float(v.grad.double().sum())); gabs.append(float(v.grad.double().abs().sum()))
    out = {'depth%d' % i: o.detach().numpy() for i, o in enumerate(outs)}
    out.update(meta=np.array([n, h, w, seed, n_resolution]), loss=np.float64(float(loss)), grad_names=np.array(names),
               grad_sum=np.array(gsum), grad_abs=np.array(gabs), grad_none=np.array(gnone),
               param_order=np.array([k for k, _ in model.decoder.named_parameters()]))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'ok  loss', float(loss), [tuple(o.shape) for o in outs])


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'multires':
        multires_case('fusionnet_multires3_2x64x96', 3, 2, 64, 96, 81)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'checkpoint':
        checkpoint_case('reference_checkpoint_tiny', 71)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'crop':
        crop_case('crop_origins_90x160')
        sys.exit(0)
    fusionnet_case('fusionnet_small_2x64x96', synth.SMALL_FUSIONNET, 2, 64, 96, 3, 'quasi_dense', True)
    fusionnet_case('fusionnet_canonical_1x64x128', synth.CANONICAL_FUSIONNET, 1, 64, 128, 0, 'sparse', False, train=False)
    fusionnet_case('fusionnet_canonical_2x96x160', synth.CANONICAL_FUSIONNET, 2, 96, 160, 1, 'quasi_dense', False)
    roi_pool_case()
    radarnet_case('radarnet_canonical_1x64x128_k3', dict(synth.CANONICAL_RADARNET, input_patch_size_image=(64, 64)),
                  1, 64, 128, 3, 2)
    s2_case('s2_compat_k6', 64, 160, (64, 64), 6, 11)
    s2_case('s2_compat_alias_k3', 32, 96, (32, 32), 3, 12, zs=[2.7, 2.2, 1.9])
    png16_case('png16_roundtrip_48x64', 48, 64, 5)
    transforms_case('transforms_5x18x26', 21)
    losses_case('losses_2x24x40', 31)
    validate_case('validate_3x20x32', 41)
    s1_case('s1_merge_64x96', 64, 96, 51)
    radarnet_loss_case('radarnet_loss_3x64x64', 61)
    checkpoint_case('reference_checkpoint_tiny', 71)
    crop_case('crop_origins_90x160')
    multires_case('fusionnet_multires3_2x64x96', 3, 2, 64, 96, 81)
    print('golden fixtures written to', OUT)
