"""The C-ABI library loads, exports every symbol include/*.h declares, and fails loudly without a GPU."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from object_slam_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(obs_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_capi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"


def test_python_binding_covers_header():
    assert set(_capi.declared_symbols()) == set(_header_symbols())


def test_version_and_keypoint_layout():
    assert b"sm_100a" in _capi.lib().obs_version()
    assert _capi.KEYPOINT_DTYPE.itemsize == 28
    assert C.sizeof(_capi.OrbParams) == 20


def test_no_cpu_fallback():
    """Without a device every compute entry point reports OBS_ERR_CUDA instead of computing on the host."""
    L = _capi.lib()
    if L.obs_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    prm = _capi.OrbParams(1000, 1.2, 8, 20, 7)
    rc = L.obs_extractor_create(C.byref(prm), 640, 480, 1, 0, C.byref(h))
    assert rc == _capi.OBS_ERR_CUDA and not h.value
    assert b"no CUDA device" in L.obs_last_error() or b"CPU" in L.obs_last_error()
    with pytest.raises(_capi.ObsError):
        from object_slam_b200.extractor import ORBextractor
        ORBextractor(1000, 1.2, 8, 20, 7)


def test_argument_validation_without_gpu():
    L = _capi.lib()
    h = C.c_void_p()
    bad = _capi.OrbParams(1000, 1.2, 99, 20, 7)
    assert L.obs_extractor_create(C.byref(bad), 640, 480, 1, 0, C.byref(h)) == _capi.OBS_ERR_INVALID
    assert L.obs_extractor_create(None, 640, 480, 1, 0, C.byref(h)) == _capi.OBS_ERR_INVALID
    n = C.c_int(5)
    # a null / empty image returns 0 keypoints like the reference's early return (ORBextractor.cc:1046)
    assert L.obs_extract(None, None, 0, 0, 0, None, None, 0, C.byref(n)) == _capi.OBS_OK and n.value == 0


def test_launch_options_without_gpu():
    """obs_set_option only flips process-wide launch switches: known names are accepted (and restored), unknown ones rejected."""
    L = _capi.lib()
    for name in (b"pdl", b"graphs", b"knn2_cta_pair"):
        assert L.obs_set_option(name, 0) == _capi.OBS_OK
        assert L.obs_set_option(name, 1) == _capi.OBS_OK
    assert L.obs_set_option(b"no_such_option", 1) == _capi.OBS_ERR_INVALID
    assert b"unknown option" in L.obs_last_error()
    assert L.obs_set_option(None, 1) == _capi.OBS_ERR_INVALID
