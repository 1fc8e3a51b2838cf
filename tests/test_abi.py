"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/rcfd.h declares (no compute calls here -- there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
    import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'rcfd.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rcfd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n


def test_binding_covers_header(lib_path):
    from rcfd import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.load()
    assert lib.rcfd_arch() == b'sm_100a'
    assert lib.rcfd_version().startswith(b'rcfd-b200')


def test_sass_is_sm100a(lib_path):
    out = subprocess.run(['cuobjdump', '-lelf', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_bad_descriptor_is_rejected_without_gpu(lib_path):
    from rcfd import _lib
    lib = _lib.load()
    d = _lib.ConvDesc()
    rc = lib.rcfd_conv2d_fwd(ctypes.byref(d), None)
    assert rc == -1 and b'conv' in lib.rcfd_last_error()


def test_row_chunk_planner_balances_waves(lib_path):
    """Host-only logic of the row-streaming kernels: chunks are chosen so that the persistent CTAs run full waves
    (e.g. 352 rows x 48 strips over 296 CTAs -> 6 chunks of 59 rows = 288 items in ONE wave, not 13 chunks = 624 items
    in 2.1 waves), every row is covered exactly once and the minimum chunk height is respected."""
    import ctypes
    from rcfd import _lib
    lib = _lib.load()

    def plan(h, cols, ctas, overhead=6, min_rows=4):
        rpc, cpc = ctypes.c_int32(), ctypes.c_int32()
        assert lib.rcfd_plan_row_chunks(h, cols, ctas, overhead, min_rows, ctypes.byref(rpc), ctypes.byref(cpc)) == 0
        return rpc.value, cpc.value

    assert plan(352, 48, 296) == (59, 6)                       # deconv0.conv, batch 8: one wave of 288 items
    assert plan(88, 16, 148) == (10, 9)                        # blocks2, batch 8: 144 items, one wave
    for h, cols, ctas in [(352, 48, 296), (176, 24, 148), (64, 2, 296), (7, 3, 148), (1000, 1, 5), (33, 200, 148)]:
        rpc, cpc = plan(h, cols, ctas)
        assert (cpc - 1) * rpc < h <= cpc * rpc                # chunks tile the rows exactly once
        assert rpc >= min(4, h) or cpc == 1
        items = cols * cpc
        waves = -(-items // ctas)
        # no other chunk count gives a cheaper schedule under the planner's own cost model
        best = min((-(-(cols * -(-h // r)) // ctas)) * (r + 6) for r in range(min(4, h), h + 1))
        assert waves * (rpc + 6) <= best
    assert lib.rcfd_plan_row_chunks(0, 1, 1, 6, 4, None, None) != 0
