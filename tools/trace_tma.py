"""In-kernel timeline of the per-tap conv engine (needs a trace build: RCFD_TRACE=1 python radar-camera-fusion-depth_b200/build.py --force).
    python tools/trace_tma.py case [case ...]       (cases of tools/bench_layers.py)
Prints, for sampled CTAs, the microseconds (clock64 / 1.965 GHz) from kernel entry to each phase."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bench_layers  # noqa: E402
from rcfd import _lib  # noqa: E402

lib = _lib.load()
names = ['entry', 'setup done', 'producer start', 'producer done', 'mma: 1st stage full', 'mma: tile0 issued',
         'mma: last tile issued', 'epi: tile0 acc ready', 'epi: tile0 drained', 'epi: tile0 stats done', 'epi: last tile done',
         'kernel end', '#tiles', 'ksteps', 'mma: 2nd round', 'mma: 3rd round', 'mma: 4th round', 'mma: 5th round', 'mma: 6th round']
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0')
for case in sys.argv[1:]:
    run, flops, byts = bench_layers.build(*bench_layers.CASES[case])
    for cold in (True, False):
        run()
        torch.cuda.synchronize()
        lib.rcfd_debug_clear_trace()
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (148 * 32))()
        (lib.rcfd_debug_read_wtrace if bench_layers.CASES[case][0] == 'wgrad' else lib.rcfd_debug_read_trace)(buf, 148 * 32)
        t = np.array(buf[:], dtype=np.int64).reshape(148, 32)
        live = [i for i in range(148) if t[i, 0] != 0 and t[i, 11] > t[i, 0]]
        print('=== %s (%s L2): %d CTAs, event time %.1f us' % (case, 'cold' if cold else 'warm', len(live), a.elapsed_time(b) * 1e3))
        for i in (live[0], live[len(live) // 2], live[-1]):
            row = t[i]
            print('  CTA %3d tiles %d ksteps %d: ' % (i, row[12], row[13]) +
                  ', '.join('%s %.2f' % (names[s], (row[s] - row[0]) / 1965.0) for s in (1, 2, 4, 14, 15, 16, 17, 5, 3, 6, 7, 8, 9, 10, 11) if row[s] != 0))
