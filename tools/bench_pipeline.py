"""End-to-end latency of one frame through both stages in memory (rcfd.bridge): camera image + K radar points ->
RadarNet stage-1 + S2 scatter -> PNG-equivalent quantisation -> FusionNet -> dense depth, 352x704, bf16."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
import fusionnet_model  # noqa: E402
import radarnet_model  # noqa: E402
from rcfd import bridge, synth  # noqa: E402

dev = torch.device('cuda:0')
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, W = 352, 704
torch.manual_seed(0)
rn = radarnet_model.RadarNetModel(device=dev, **synth.CANONICAL_RADARNET)
rn.set_precision('bf16')
rn.eval()
fn = fusionnet_model.FusionNetModel(device=dev, **synth.CANONICAL_FUSIONNET)
fn.set_precision('bf16')
fn.eval()
image = torch.rand(1, 3, H, W, device=dev)
pts = synth.radar_points(k, H, W, 0).to(dev)


def timed(fn_, reps=10):
    for _ in range(3):
        fn_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn_()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


with torch.no_grad():
    t_all = timed(lambda: bridge.image_and_radar_to_depth(rn, fn, image, pts))
    t_s1 = timed(lambda: bridge.radar_to_input_depth(rn, image, pts))
    inp = bridge.radar_to_input_depth(rn, image, pts)
    t_s2 = timed(lambda: fn.forward(image, inp))
    t_s2g = timed(lambda: fn.forward_graphed(image, inp))
print('frame 352x704, %d radar points, bf16: image+points -> depth %.2f ms (%.0f frames/s); stage 1 + S2 + bridge %.2f ms; '
      'FusionNet %.2f ms eager, %.2f ms graphed' % (k, t_all, 1e3 / t_all, t_s1, t_s2, t_s2g))
