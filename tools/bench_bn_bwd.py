"""BatchNorm backward of the small maps of a batch-8 step (levels <= 22x44): one launch (rcfd_bn_act_bwd_fused) against
leaky_bwd + reduce + apply, timed the way the step graph runs them -- 20 dependent calls captured in a CUDA graph, L2 warm.

    python tools/bench_bn_bwd.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
from rcfd import ops  # noqa: E402

dev = torch.device('cuda:0')
REPS = 20
SHAPES = [('6x11   x256', 528, 256), ('11x22  x256', 1936, 256), ('22x44  x256', 7744, 256), ('22x44  x128', 7744, 128),
          ('44x88  x128', 30976, 128), ('44x88  x64', 30976, 64), ('88x176 x64', 123904, 64), ('352x704x32', 1982464, 32)]


def time_graph(fn, pool):
    """fn(i): the i-th of REPS dependent calls; pool: the zeroed scratch they take their slices from (zeroed once per
    replay, like the step's pool)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pool.zero_()
        fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            pool.zero_()
            for i in range(REPS):
                fn(i)
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / REPS * 1e3


if __name__ == '__main__':
    print('%-14s %5s | %9s %12s' % ('map', 'post', 'fused us', '3-kernel us'))
    for name, pixels, c in SHAPES:
        y = torch.randn(8, 1, pixels // 8, c, device=dev).bfloat16()
        dz = torch.randn_like(y)
        z = torch.randn_like(y)
        scale, shift, mean = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev), torch.randn(c, device=dev)
        invstd = torch.rand(c, device=dev) + 0.5
        dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
        pool = torch.zeros(REPS, 2 * c, device=dev, dtype=torch.float64)
        for post in (None, z):
            res = []
            for limit in (1 << 20, 0):
                ops.BN_BWD_FUSED_MAX_PIXELS = limit
                if limit and pixels > 32768:
                    res.append(float('nan'))
                    continue
                res.append(time_graph(lambda i: ops.bn_act_bwd(dz, y, scale, shift, mean, invstd, ops.ACT_LEAKY, dg, db,
                                                               sums=pool[i], post_z=post), pool))
            print('%-14s %5s | %9.2f %12.2f' % (name, 'yes' if post is not None else 'no', res[0], res[1]), flush=True)
