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


class FrameParams(C.Structure):
    _fields_ = [("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("mbf", C.c_float), ("mb", C.c_float), ("nlevels", C.c_int32), ("scale_factors", C.c_float * 12)]


class FrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("descriptors", C.c_void_p), ("u_right", C.c_void_p)]


class MapPointView(C.Structure):
    _fields_ = [("n", C.c_int32), ("per_frame", C.c_int32), ("in_view", C.c_void_p), ("proj_x", C.c_void_p),
                ("proj_y", C.c_void_p), ("proj_xr", C.c_void_p), ("scale_level", C.c_void_p), ("view_cos", C.c_void_p),
                ("descriptors", C.c_void_p), ("observations", C.c_void_p)]


class LastFrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("per_frame", C.c_int32), ("has_point", C.c_void_p), ("world_pos", C.c_void_p),
                ("octave", C.c_void_p), ("angle", C.c_void_p), ("descriptors", C.c_void_p), ("observations", C.c_void_p),
                ("tcw_last", C.c_void_p), ("tcw_current", C.c_void_p)]


class KeyFramePointsView(C.Structure):
    _fields_ = [("n", C.c_int32), ("per_frame", C.c_int32), ("valid", C.c_void_p), ("world_pos", C.c_void_p),
                ("min_distance", C.c_void_p), ("max_distance", C.c_void_p), ("max_distance_raw", C.c_void_p),
                ("normal", C.c_void_p), ("angle", C.c_void_p), ("descriptors", C.c_void_p), ("tcw", C.c_void_p)]


class BowSide(C.Structure):
    _fields_ = [("cap", C.c_int32), ("node_cap", C.c_int32), ("n", C.c_void_p), ("descriptors", C.c_void_p), ("keys_un", C.c_void_p),
                ("valid", C.c_void_p), ("u_right", C.c_void_p), ("n_nodes", C.c_void_p), ("node_id", C.c_void_p),
                ("node_start", C.c_void_p), ("node_idx", C.c_void_p)]


class StereoIO(C.Structure):
    _fields_ = [("left", C.c_void_p), ("right", C.c_void_p),
                ("kp_left", C.c_void_p), ("desc_left", C.c_void_p), ("n_left", C.c_void_p),
                ("kp_right", C.c_void_p), ("desc_right", C.c_void_p), ("n_right", C.c_void_p),
                ("u_right", C.c_void_p), ("depth", C.c_void_p)]


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
    "obs_set_option": (C.c_int, [C.c_char_p, C.c_int]),
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
    "obs_extract_batch_submit": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, _vp, _vp, C.c_int, _vp]),
    "obs_extract_batch_wait": (C.c_int, [_vp]),
    "obs_stereo_frames_submit": (C.c_int, [_vp, _vp, C.POINTER(StereoIO), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_float]),
    "obs_stereo_frames_wait": (C.c_int, [_vp, _vp]),
    "obs_stereo_frames": (C.c_int, [_vp, _vp, C.POINTER(StereoIO), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_float]),
    "obs_stereo_match_device": (C.c_int, [_vp, _vp, C.c_float, C.c_float, C.c_float, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "obs_gray_from_color": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp, C.c_size_t, C.c_size_t, _vp]),
    "obs_depth_to_float": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_float, _vp, C.c_size_t, C.c_size_t, _vp]),
    "obs_stereo_from_rgbd": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_float, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "obs_matcher_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "obs_matcher_destroy": (C.c_int, [_vp]),
    "obs_matcher_stream": (_vp, [_vp]),
    "obs_matcher_sync": (C.c_int, [_vp]),
    "obs_matcher_last_rounds": (C.c_int, [_vp, _vp, C.c_int]),
    "obs_frame_set_create": (C.c_int, [_vp, C.POINTER(FrameParams), C.c_int, C.c_int, C.POINTER(_vp)]),
    "obs_frame_set_destroy": (C.c_int, [_vp]),
    "obs_frame_set_upload": (C.c_int, [_vp, C.POINTER(FrameView), C.c_int]),
    "obs_frame_set_from_extractor": (C.c_int, [_vp, _vp, _vp]),
    "obs_frame_set_count": (C.c_int, [_vp]),
    "obs_frame_set_grid": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "obs_search_by_projection": (C.c_int, [_vp, _vp, C.POINTER(MapPointView), C.c_float, C.c_float, _vp, _vp, _vp]),
    "obs_search_by_projection_last": (C.c_int, [_vp, _vp, C.POINTER(LastFrameView), C.c_float, C.c_int, C.c_int, _vp, _vp, _vp]),
    "obs_search_by_projection_keyframe": (C.c_int, [_vp, _vp, C.POINTER(KeyFramePointsView), C.c_float, C.c_int, C.c_int, _vp, _vp, _vp]),
    "obs_search_by_projection_sim3": (C.c_int, [_vp, _vp, C.POINTER(KeyFramePointsView), C.c_int, _vp, _vp, _vp]),
    "obs_fuse_search": (C.c_int, [_vp, _vp, C.POINTER(KeyFramePointsView), _vp, C.c_float, C.c_int, _vp, _vp]),
    "obs_search_by_sim3": (C.c_int, [_vp, _vp, _vp, C.POINTER(KeyFramePointsView), C.POINTER(KeyFramePointsView), _vp, _vp, C.c_float, _vp, _vp]),
    "obs_search_for_initialization": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int, _vp]),
    "obs_compute_three_maxima": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "obs_descriptor_distance": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "obs_comm_nccl_version": (C.c_int, []),
    "obs_comm_unique_id": (C.c_int, [_vp]),
    "obs_comm_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "obs_comm_destroy": (C.c_int, [_vp]),
    "obs_comm_allgather": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_int, _vp]),
    "obs_comm_wait": (C.c_int, [_vp, C.c_int, _vp]),
    "obs_microbench_imma": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "obs_microbench_pipes": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "obs_microbench_popc": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "obs_search_by_bow": (C.c_int, [_vp, C.POINTER(BowSide), C.POINTER(BowSide), C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp]),
    "obs_search_for_triangulation": (C.c_int, [_vp, C.POINTER(BowSide), C.POINTER(BowSide), C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "obs_distinctive_descriptors": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "obs_assign_keypoints_to_masks": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_float, C.c_int, _vp, _vp, _vp, _vp]),
    "obs_hsv_histograms": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp]),
    "obs_undistort_keypoints": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vp, C.c_int, _vp]),
    "obs_undistort_points": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vp, C.c_int, _vp]),
    "obs_distance_transform": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp]),
    "obs_matcher_set_knn2_engine": (C.c_int, [_vp, C.c_int]),
    "obs_hamming_knn2": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_float, _vp, _vp, _vp]),
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
    """Address of a numpy array; ints pass through as raw (host or device) addresses."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def addr(a):
    """The same as an integer (0 for None), for ctypes.Structure fields."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


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
