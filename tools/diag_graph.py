"""Per-parameter comparison of first-step gradients: eager vs eager (run-to-run determinism) and
graphed vs eager (tools/diag_graph.py [n h w]).  Diagnostic only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'radar-camera-fusion-depth_b200'), os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import fusionnet_model  # noqa: E402
import net_utils  # noqa: E402
from rcfd import optim, synth  # noqa: E402
from helpers import synth_fusionnet_state  # noqa: E402

DEV = torch.device('cuda:0')
n, h, w = [int(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 else (2, 96, 160)
cfg = synth.CANONICAL_FUSIONNET
p0 = synth_fusionnet_state(cfg, 7)
image, depth = synth.fusionnet_inputs(n, h, w, 7, 'quasi_dense')
gt, lidar = synth.training_targets(n, h, w, 7)
batch = [t.to(DEV) for t in (image, depth, gt, lidar)]
outlier = net_utils.OutlierRemoval(7, 1.5)


def run(graphed, steps=1):
    m = fusionnet_model.FusionNetModel(device=DEV, **cfg)
    m.encoder.load_state_dict({k[len('encoder.'):]: v for k, v in p0.items() if k.startswith('encoder.')})
    m.decoder.load_state_dict({k[len('decoder.'):]: v for k, v in p0.items() if k.startswith('decoder.')})
    m.set_precision('bf16')
    m.train()
    opt = optim.FusedAdam(m.parameters(), lr=1e-3, eps=float(os.environ.get('ADAM_EPS', '1e-8')))
    names = [('encoder.' + k) for k, _ in m.encoder.named_parameters()] + [('decoder.' + k) for k, _ in m.decoder.named_parameters()]
    out = []
    for _ in range(steps):
        im, dp, g, l = batch
        if graphed:
            loss = m.train_step_graphed(im, dp, g, l, opt, 2.0, outlier_removal=outlier)
        else:
            d = m.forward(im, dp)
            loss, _ = m.compute_loss(im, d, outlier.remove_outliers(g), l, 'l1', 0.0, -1, None, 2.0)
            loss.backward()
            opt.step()
        out.append((float(loss), {k: p.grad.detach().clone() for k, p in zip(names, m.parameters())},
                    {k: p.detach().clone() for k, p in zip(names, m.parameters())}))
    return out


def compare(tag, a, b):
    print('==', tag, 'loss', a[0], b[0])
    rows = []
    for k in a[1]:
        x, y = a[1][k].double(), b[1][k].double()
        den = float(y.abs().max()) + 1e-30
        rows.append((float((x - y).abs().max()) / den, k, tuple(x.shape), den))
    rows.sort(reverse=True)
    for r in rows[:12]:
        print('  %.3e  %-50s %s  max|ref| %.3e' % r)


e1 = run(False, 2)
e2 = run(False, 2)
g1 = run(True, 2)
compare('eager vs eager, step 1', e1[0], e2[0])
compare('graph vs eager, step 1', g1[0], e1[0])
compare('eager vs eager, step 2', e1[1], e2[1])
compare('graph vs eager, step 2', g1[1], e1[1])


def compare_params(tag, a, b):
    nbad, ntot, worst = 0, 0, 0.0
    for k in a[2]:
        d = (a[2][k] - b[2][k]).abs()
        bad = d > 1e-4
        nbad += int(bad.sum())
        ntot += d.numel()
        if bool(bad.any()):
            gmax = float(b[1][k].abs().max())
            worst = max(worst, float((b[1][k].abs()[bad]).max()) / (gmax + 1e-30))
    print('==', tag, 'params differing by > 1e-4 after the step: %d of %d; largest |g|/max|g| among them %.3e' % (nbad, ntot, worst))


compare_params('eager vs eager step 1', e1[0], e2[0])
compare_params('graph vs eager step 1', g1[0], e1[0])
