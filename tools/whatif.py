"""What-if timing experiments (measurement tool, WRONG RESULTS by construction, never a bench number): run bench.py with
selected C-ABI entry points turned into no-ops, to see how much of the step's WALL time (multi-stream, CUDA graph) each
kernel family is responsible for -- the serialised per-kernel times of rcfd/census.py do not show what is on the critical
path.

    python tools/whatif.py rcfd_conv2d_wgrad,rcfd_unpack_conv_wgrad -- --no-cpu --no-also --no-census --no-gpu-baseline
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))

skip = set(filter(None, sys.argv[1].split(',')))
rest = sys.argv[3:] if len(sys.argv) > 2 and sys.argv[2] == '--' else sys.argv[2:]

from rcfd import _lib  # noqa: E402

_real = _lib.call


def call(name, *args):
    if name in skip:
        _lib.launch_count += 1
        return None
    return _real(name, *args)


_lib.call = call
sys.argv = ['bench.py'] + rest
import bench  # noqa: E402

print('WHAT-IF (wrong results): skipping', sorted(skip), file=sys.stderr)
bench.main()
