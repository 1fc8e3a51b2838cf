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
