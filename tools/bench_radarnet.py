"""RadarNet stage-1 forward + S2 scatter timing (BASELINE configs[2]: 352x704 image, 352x288 patches, K = 64 points)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
import radarnet_model  # noqa: E402
import radarnet_main  # noqa: E402
from rcfd import synth  # noqa: E402

dev = torch.device('cuda:0')
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 16
k = 64
H, W = 352, 704
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
torch.manual_seed(0)
m = radarnet_model.RadarNetModel(device=dev, **synth.CANONICAL_RADARNET)
m.set_precision(precision)
m.eval()
pad = 288 // 2
images = [torch.rand(1, 3, H, W, device=dev) for _ in range(n_img)]
pts, boxes = [], []
for b in range(n_img):
    pt = synth.radar_points(k, H, W, b)
    pt[:, 0] += pad
    pts.append(pt.to(dev))
    boxes.append([torch.stack([pt[:, 0] - pad, torch.zeros(k), pt[:, 0] + pad, torch.full((k,), float(H))], 1).to(dev)])


def step():
    outs = []
    with torch.no_grad():
        for b in range(n_img):            # the reference's entry point is per image (radarnet_main.forward)
            outs.append(radarnet_main.forward(m, images[b], pts[b], boxes[b], device=dev))
    return outs


for _ in range(2):
    out = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    out = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
d, r = out[0]
print('radarnet stage-1 (%s): %d images x %d points, %.2f ms per batch, %.1f images/s, %.0f point-columns/s; '
      'response>0 pixels in image 0: %d, depth dtype %s'
      % (precision, n_img, k, ms, n_img / ms * 1e3, n_img * k / ms * 1e3, int((r > 0).sum()), d.dtype))
gflop = (14.35 + 8.20 * k) * n_img
print('algorithmic %.0f GFLOP per batch -> %.1f TFLOP/s' % (gflop, gflop / ms))
