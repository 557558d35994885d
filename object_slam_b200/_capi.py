"""ctypes binding of libobslam_b200.so (the C ABI declared in include/obslam_b200.h).

The library is built in-tree by ``object_slam_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: if the shared object is missing or cannot be loaded, importing a symbol
from here raises, and every compute call fails with the library's own error when no sm_100
device is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libobslam_b200.so")

OBS_OK, OBS_ERR_INVALID, OBS_ERR_CUDA, OBS_ERR_CAPACITY, OBS_ERR_STATE = range(5)

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")])
assert KEYPOINT_DTYPE.itemsize == 28


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class ObsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"obslam_b200 error {code}: {msg}")
        self.code = code


_vp = C.c_void_p
_PROTOS = {
    # name: (restype, argtypes)
    "obs_last_error": (C.c_char_p, []),
    "obs_version": (C.c_char_p, []),
    "obs_device_count": (C.c_int, []),
    "obs_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "obs_host_free": (C.c_int, [_vp]),
    "obs_extractor_create": (C.c_int, [C.POINTER(OrbParams), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "obs_extractor_destroy": (C.c_int, [_vp]),
    "obs_extractor_levels": (C.c_int, [_vp]),
    "obs_extractor_max_keypoints": (C.c_int, [_vp]),
    "obs_extractor_tables": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "obs_extract": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_size_t, _vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "obs_extract_batch": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_size_t, _vp, _vp, C.c_int, _vp]),
    "obs_extract_batch_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp]),
    "obs_extractor_fetch": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "obs_extractor_fetch_counts": (C.c_int, [_vp, _vp]),
    "obs_extractor_results_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "obs_extractor_get_level": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "obs_extractor_get_candidates": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, C.POINTER(C.c_int)]),
    "obs_extractor_get_selected": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, C.POINTER(C.c_int)]),
    "obs_extractor_set_profiling": (C.c_int, [_vp, C.c_int]),
    "obs_extractor_stage_ms": (C.c_int, [_vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "obs_stereo_match": (C.c_int, [_vp, _vp, C.c_float, C.c_float, C.c_float, _vp, _vp, C.c_int]),
    "obs_stereo_match_device": (C.c_int, [_vp, _vp, C.c_float, C.c_float, C.c_float, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C object_slam_b200/csrc` "
                "(or __graft_entry__.build()); object_slam_b200 has no non-CUDA path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)      # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def declared_symbols():
    return sorted(_PROTOS)


def check(rc):
    if rc != OBS_OK:
        raise ObsError(rc, lib().obs_last_error().decode("utf-8", "replace"))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _PinnedOwner:
    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            lib().obs_host_free(self.p)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (obs_host_alloc): moved by DMA without a staging copy."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = C.c_void_p()
    check(lib().obs_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    buf._owner = _PinnedOwner(p)      # numpy keeps `buf` alive through .base; the allocation dies with it
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
