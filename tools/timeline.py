"""Kernel timeline of ONE graphed training step (torch.profiler / CUPTI timestamps; analysis only, never a bench number).
Prints per-stream busy time, the step span, and the kernels on the busiest stream with the gaps between them."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fusionnet_model  # noqa: E402
import net_utils  # noqa: E402
from rcfd import optim, synth  # noqa: E402

dev = torch.device('cuda:0')
mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
m = fusionnet_model.FusionNetModel(device=dev, **synth.CANONICAL_FUSIONNET)
m.set_precision('bf16')
data = [t.to(dev) for t in bench.synthetic_batch(batch, 0)]
if mode == 'train':
    m.train()
    opt = optim.FusedAdam(m.parameters(), lr=1e-3)
    outlier = net_utils.OutlierRemoval(7, 1.5)
    step = lambda: m.train_step_graphed(data[0], data[1], data[2], data[3], opt, 2.0, outlier_removal=outlier)
else:
    m.eval()
    step = lambda: m.forward_graphed(data[0], data[1])
with torch.no_grad():
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() >= 0]
ks = []
for e in ev:
    name = e.name
    if 'Memcpy' in name or 'Memset' in name or 'memcpy' in name or 'memset' in name:
        kind = 'mem'
    else:
        kind = 'k'
    ks.append((e.time_range.start, e.time_range.end, getattr(e, 'device_index', 0), name, kind, getattr(e, 'stream', None)))
ks.sort()
if not ks:
    raise SystemExit('no CUDA events captured')
t0 = ks[0][0]
t1 = max(k[1] for k in ks)
print('events %d, span %.3f ms' % (len(ks), (t1 - t0) / 1e3))
# torch's FunctionEvent has no stream id: recover lanes greedily (a kernel goes to the first lane that is free)
lanes = []
lane_of = []
for k in ks:
    for i, end in enumerate(lanes):
        if end <= k[0]:
            lanes[i] = k[1]
            lane_of.append(i)
            break
    else:
        lanes.append(k[1])
        lane_of.append(len(lanes) - 1)
busy = defaultdict(float)
for k, l in zip(ks, lane_of):
    busy[l] += (k[1] - k[0]) / 1e3
print('concurrency lanes (greedy):', {l: round(b, 3) for l, b in sorted(busy.items())})
# time covered by at least one kernel, and average concurrency
pts = sorted([(k[0], 1) for k in ks] + [(k[1], -1) for k in ks])
cov, depth, last, area = 0.0, 0, pts[0][0], 0.0
hist = defaultdict(float)
for t, d in pts:
    if depth > 0:
        cov += t - last
    hist[depth] += t - last
    area += depth * (t - last)
    depth += d
    last = t
print('covered %.3f ms, idle %.3f ms, mean concurrency while busy %.2f' % (cov / 1e3, (t1 - t0 - cov) / 1e3, area / max(cov, 1)))
print('time at concurrency d (ms):', {d: round(v / 1e3, 3) for d, v in sorted(hist.items())})
# the long kernels and when they ran
print('--- kernels >= 40 us (start ms, dur us, concurrent lanes at start)')
for k, l in zip(ks, lane_of):
    if k[1] - k[0] >= 40:
        print('%8.3f %7.1f  lane %d  %s' % ((k[0] - t0) / 1e3, k[1] - k[0], l, k[3][:70]))
if '--all' in sys.argv:
    print('--- all')
    for k, l in zip(ks, lane_of):
        print('%8.3f %7.1f  lane %d  %s' % ((k[0] - t0) / 1e3, k[1] - k[0], l, k[3][:60]))
