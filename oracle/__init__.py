"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``object_slam_b200/``
does: the product path fails loudly when its CUDA library is missing rather than fall
back to any of this.

Two libraries:
  * ``liborb_oracle.so``  -- restatement of the reference algorithm (orb_oracle.cpp,
    match_oracle.cpp), each function citing the reference file:line it follows.
  * ``_ref/libref_orbextractor.so`` -- the reference's own ``src/ORBextractor.cc`` compiled
    unmodified against ``cvshim`` (only buildable where /root/reference exists; the built
    file travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liborb_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libref_orbextractor.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")])
assert KEYPOINT_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or _stale():
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])


def _stale():
    t = os.path.getmtime(LIB_PATH)
    for f in ("orb_oracle.cpp", "match_oracle.cpp", "orc_primitives.h"):
        if os.path.getmtime(os.path.join(HERE, f)) > t:
            return True
    return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_extractor_create.restype = C.c_void_p
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_destroy.argtypes = [C.c_void_p]
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.orc_get_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_get_level_dims.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
        L.orc_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_get_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.orc_gaussian7x7_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_fast9_16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast9_16_simple.argtypes = L.orc_fast9_16.argtypes
        L.orc_fast_score_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast_atan2_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_sincosf_model_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_sincosf_libm_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_sincosf_sweep.restype = C.c_long
        L.orc_sincosf_sweep.argtypes = [C.c_float, C.c_float, C.c_long]
        L.orc_pattern.restype = C.POINTER(C.c_int * 1024)
        _bind_match(L)
        _lib = L
    return _lib


def _bind_match(L):
    L.orc_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_frame_create.restype = C.c_void_p
    L.orc_frame_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_float] * 4
    L.orc_frame_destroy.argtypes = [C.c_void_p]
    L.orc_frame_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_features_in_area.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orc_compute_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_search_by_projection_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection_last.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_for_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
    L.orc_hamming_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection_keyframe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_by_bow.argtypes = [C.c_void_p] * 14 + [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_for_triangulation.argtypes = [C.c_void_p] * 17 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_fuse_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_void_p, C.c_void_p]
    L.orc_search_by_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float] + [C.c_void_p] * 5 + ([C.c_int] + [C.c_void_p] * 6) * 2 + [C.c_float, C.c_void_p]
    L.orc_assign_keypoints_to_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_hsv_from_bgr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_hsv_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_distance_transform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_logf.restype = C.c_float
    L.orc_logf.argtypes = [C.c_float]
    L.orc_norm3.restype = C.c_float
    L.orc_norm3.argtypes = [C.c_void_p]
    L.orc_predict_scale.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    L.orc_project_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_minus_rt_t.argtypes = [C.c_void_p, C.c_void_p]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


# ----------------------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = _img(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dw)
    return dst


def gaussian7x7(src):
    src = _img(src)
    dst = np.empty_like(src)
    lib().orc_gaussian7x7_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dst.strides[0])
    return dst


def fast9_16(img, threshold, nms=True, simple=False):
    img = _img(img)
    cap = img.size
    out = np.empty((max(cap, 1), 3), np.int32)
    fn = lib().orc_fast9_16_simple if simple else lib().orc_fast9_16
    n = fn(_ptr(img), img.shape[1], img.shape[0], img.strides[0], int(threshold), int(nms), _ptr(out), cap)
    return out[:n].copy()


def fast_score_map(img):
    img = _img(img)
    out = np.empty_like(img)
    lib().orc_fast_score_map(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out))
    return out


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().orc_fast_atan2_n(_ptr(y), _ptr(x), _ptr(out), y.size)
    return out


def sincosf(a, model=True):
    a = np.ascontiguousarray(a, np.float32)
    s = np.empty_like(a)
    c = np.empty_like(a)
    (lib().orc_sincosf_model_n if model else lib().orc_sincosf_libm_n)(_ptr(a), _ptr(s), _ptr(c), a.size)
    return s, c


def pattern():
    return np.array(lib().orc_pattern().contents, dtype=np.int32).reshape(256, 4)


# ----------------------------------------------------------------------------- extractor
class OracleExtractor:
    """Restated ORBextractor (orb_oracle.cpp); keeps every intermediate of the last call."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self._h = lib().orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_extractor_destroy(self._h)
            self._h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        fpl = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        lib().orc_get_tables(self._h, _ptr(sc), _ptr(isc), _ptr(s2), _ptr(is2), _ptr(fpl), _ptr(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, features_per_level=fpl, umax=umax)

    def __call__(self, img):
        img = _img(img)
        n = lib().orc_extract(self._h, _ptr(img), img.shape[1], img.shape[0], img.strides[0])
        kps = np.empty(n, KEYPOINT_DTYPE)
        desc = np.empty((n, 32), np.uint8)
        lib().orc_get_keypoints(self._h, _ptr(kps), _ptr(desc), n)
        return kps, desc

    def level(self, level, blurred=False):
        w, h = C.c_int32(), C.c_int32()
        lib().orc_get_level_dims(self._h, level, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        got = lib().orc_get_level(self._h, int(blurred), level, _ptr(out))
        return out if got else None

    def level_keypoints(self, level, selected):
        n = lib().orc_get_level_keypoints(self._h, int(selected), level, None, 0)
        out = np.empty(n, KEYPOINT_DTYPE)
        lib().orc_get_level_keypoints(self._h, int(selected), level, _ptr(out), n)
        return out


# ----------------------------------------------------------------------------- reference
_ref = None


def ref_available():
    return os.path.exists(REF_PATH)


def ref_lib():
    global _ref
    if _ref is None:
        if not ref_available():
            build(force=True)
        R = C.CDLL(REF_PATH)
        R.ref_extractor_create.restype = C.c_void_p
        R.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        R.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        R.ref_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, _i32p, _i32p]
        R.ref_get_scale_factors.argtypes = [C.c_void_p, C.c_void_p]
        if hasattr(R, "ref_extract_many"):
            R.ref_extract_many.restype = C.c_double
            R.ref_extract_many.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _ref = R
    return _ref


class ReferenceExtractor:
    """The reference's own ORB_SLAM2::ORBextractor (unmodified source, cv shim, bump allocator).

    Create and call on one thread (the handle's tables live in that thread's arena)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.cap = nfeatures * 2 + 1024
        self._h = ref_lib().ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __call__(self, img):
        img = _img(img)
        kps = np.empty(self.cap, KEYPOINT_DTYPE)
        desc = np.empty((self.cap, 32), np.uint8)
        n = ref_lib().ref_extract(self._h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), _ptr(desc), self.cap)
        assert n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, level):
        w, h = C.c_int32(), C.c_int32()
        ref_lib().ref_get_level(self._h, level, None, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        ref_lib().ref_get_level(self._h, level, _ptr(out), C.byref(w), C.byref(h))
        return out


# ----------------------------------------------------------------------------- matching
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_ptr(a), _ptr(b))


def stereo_match(kpsL, descL, kpsR, descR, pyrL, pyrR, scale, inv_scale, mbf, minD, maxD):
    """Frame::ComputeStereoMatches restated (match_oracle.cpp).  pyrL/pyrR: lists of border-less
    level images.  Returns (uRight, depth, sad) over the left keypoints (sad = -1 where the
    sub-pixel stage did not accept the match; it is not reset by the median cut)."""
    kpsL = np.ascontiguousarray(kpsL); kpsR = np.ascontiguousarray(kpsR)
    descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
    pl = [_img(p) for p in pyrL]
    pr = [_img(p) for p in pyrR]
    n = len(pl)
    arrL = (C.c_void_p * n)(*[p.ctypes.data for p in pl])
    arrR = (C.c_void_p * n)(*[p.ctypes.data for p in pr])
    lw = np.array([p.shape[1] for p in pl], np.int32)
    lh = np.array([p.shape[0] for p in pl], np.int32)
    scale = np.ascontiguousarray(scale, np.float32)
    inv_scale = np.ascontiguousarray(inv_scale, np.float32)
    nL = len(kpsL)
    ur = np.empty(max(nL, 1), np.float32); dp = np.empty(max(nL, 1), np.float32); sad = np.empty(max(nL, 1), np.int32)
    lib().orc_stereo_match(_ptr(kpsL), _ptr(descL), nL, _ptr(kpsR), _ptr(descR), len(kpsR),
                           arrL, arrR, _ptr(lw), _ptr(lh), n, _ptr(scale), _ptr(inv_scale),
                           float(mbf), float(minD), float(maxD), _ptr(ur), _ptr(dp), _ptr(sad))
    return ur[:nL], dp[:nL], sad[:nL]


class OracleFrame:
    """Flat view of a Frame for the matchers: mvKeysUn, mDescriptors, mvuRight, image bounds and the
    64x48 grid (Frame.cc:455-470, :567-632), restated in match_oracle.cpp."""

    def __init__(self, keys_un, desc, u_right=None, bounds=None):
        self.keys = np.ascontiguousarray(keys_un, KEYPOINT_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.n = len(self.keys)
        assert len(self.desc) == self.n
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.bounds = tuple(float(b) for b in bounds)          # (minX, maxX, minY, maxY)
        self._h = lib().orc_frame_create(_ptr(self.keys), _ptr(self.desc),
                                         None if self.u_right is None else _ptr(self.u_right), self.n, *self.bounds)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_frame_destroy(self._h)
            self._h = None

    def grid(self):
        start = np.empty(64 * 48 + 1, np.int32)
        idx = np.empty(max(self.n, 1), np.int32)
        lib().orc_frame_grid(self._h, _ptr(start), _ptr(idx))
        return start, idx[:start[-1]]

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(max(self.n, 1), np.int32)
        n = lib().orc_features_in_area(self._h, x, y, r, min_level, max_level, _ptr(out), len(out))
        return out[:n].copy()


def compute_three_maxima(sizes):
    sizes = np.ascontiguousarray(sizes, np.int32)
    ind = np.empty(3, np.int32)
    lib().orc_compute_three_maxima(_ptr(sizes), len(sizes), _ptr(ind))
    return tuple(int(i) for i in ind)


def _state(frame, kp_obs):
    obs = np.zeros(max(frame.n, 1), np.int32)
    if kp_obs is not None:
        obs[:frame.n] = kp_obs
    return obs, np.full(max(frame.n, 1), -1, np.int32)


def search_by_projection_map(frame, scale, in_view, proj_x, proj_y, proj_xr, level, view_cos, mp_desc, mp_obs,
                             th, nnratio, kp_obs=None):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, th).  Returns (nmatches, kp_match, kp_obs_out)."""
    scale = np.ascontiguousarray(scale, np.float32)
    in_view = np.ascontiguousarray(in_view, np.uint8)
    f = [np.ascontiguousarray(a, np.float32) for a in (proj_x, proj_y, proj_xr)]
    level = np.ascontiguousarray(level, np.int32)
    view_cos = np.ascontiguousarray(view_cos, np.float32)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    mp_obs = np.ascontiguousarray(mp_obs, np.int32)
    obs, match = _state(frame, kp_obs)
    n = lib().orc_search_by_projection_map(frame._h, _ptr(scale), len(in_view), _ptr(in_view), _ptr(f[0]), _ptr(f[1]), _ptr(f[2]),
                                           _ptr(level), _ptr(view_cos), _ptr(mp_desc), _ptr(mp_obs), th, nnratio,
                                           _ptr(obs), _ptr(match))
    return n, match[:frame.n], obs[:frame.n]


def search_by_projection_last(frame, scale, cam, tcw_cur, tcw_last, last_has_point, last_pos, last_octave, last_angle,
                              mp_desc, mp_obs, th, mono, check_ori=True, kp_obs=None):
    """ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono).  cam = (fx, fy, cx, cy, mbf, mb);
    tcw_* = 3x4 row-major [R|t]."""
    scale = np.ascontiguousarray(scale, np.float32)
    cam = np.ascontiguousarray(cam, np.float32)
    tc = np.ascontiguousarray(tcw_cur, np.float32).reshape(12)
    tl = np.ascontiguousarray(tcw_last, np.float32).reshape(12)
    has = np.ascontiguousarray(last_has_point, np.uint8)
    pos = np.ascontiguousarray(last_pos, np.float32).reshape(-1, 3)
    octv = np.ascontiguousarray(last_octave, np.int32)
    ang = np.ascontiguousarray(last_angle, np.float32)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    mp_obs = np.ascontiguousarray(mp_obs, np.int32)
    obs, match = _state(frame, kp_obs)
    n = lib().orc_search_by_projection_last(frame._h, _ptr(scale), _ptr(cam), _ptr(tc), _ptr(tl), len(has), _ptr(has), _ptr(pos),
                                            _ptr(octv), _ptr(ang), _ptr(mp_desc), _ptr(mp_obs), th, int(mono), int(check_ori),
                                            _ptr(obs), _ptr(match))
    return n, match[:frame.n], obs[:frame.n]


def search_for_initialization(f1, f2, prev_matched, window, nnratio, check_ori=True):
    """ORBmatcher::SearchForInitialization.  Returns (nmatches, matches12, prev_matched_out)."""
    pm = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.empty(max(f1.n, 1), np.int32)
    n = lib().orc_search_for_initialization(f1._h, f2._h, _ptr(pm), _ptr(m12), int(window), nnratio, int(check_ori))
    return n, m12[:f1.n], pm


def hamming_knn2(q, db, th_low=50, nnratio=0.6):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    db = np.ascontiguousarray(db, np.uint8).reshape(-1, 32)
    bi = np.empty(max(len(q), 1), np.int32); bd = np.empty_like(bi); sd = np.empty_like(bi)
    lib().orc_hamming_knn2(_ptr(q), len(q), _ptr(db), len(db), th_low, nnratio, _ptr(bi), _ptr(bd), _ptr(sd))
    return bi[:len(q)], bd[:len(q)], sd[:len(q)]


def project_points(tcw, pos):
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(12)
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    out = np.empty_like(pos)
    lib().orc_project_points(_ptr(tcw), _ptr(pos), len(pos), _ptr(out))
    return out


def minus_rt_t(tcw):
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(12)
    out = np.empty(3, np.float32)
    lib().orc_minus_rt_t(_ptr(tcw), _ptr(out))
    return out


def logf(x):
    return float(lib().orc_logf(float(x)))


def norm3(v):
    v = np.ascontiguousarray(v, np.float32)
    return float(lib().orc_norm3(_ptr(v)))


def _kf_points(pts, key):
    return [np.ascontiguousarray(pts[k], t) for k, t in (("valid", np.uint8), ("world_pos", np.float32), ("min_distance", np.float32),
                                                          ("max_distance", np.float32), ("max_distance_raw", np.float32), (key, np.float32),
                                                          ("descriptors", np.uint8))]


def search_by_projection_keyframe(frame, scale, cam, tcw, pts, th, orb_dist, check_ori=True, kp_taken=None):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist).  pts: dict(valid, world_pos, min_distance,
    max_distance, max_distance_raw, angle, descriptors).  Returns (nmatches, kp_match)."""
    scale = np.ascontiguousarray(scale, np.float32)
    cam = np.ascontiguousarray(cam, np.float32)
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(12)
    a = _kf_points(pts, "angle")
    taken, match = _state(frame, kp_taken)
    n = lib().orc_search_by_projection_keyframe(frame._h, _ptr(scale), len(scale), logf(scale[1]), _ptr(cam), _ptr(tcw), len(a[0]),
                                                *[_ptr(x) for x in a], th, int(orb_dist), int(check_ori), _ptr(taken), _ptr(match))
    return n, match[:frame.n]


def search_by_projection_sim3(frame, scale, cam, tcw, pts, th, kp_taken=None):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) after the decomposition of Scw."""
    scale = np.ascontiguousarray(scale, np.float32)
    cam = np.ascontiguousarray(cam, np.float32)
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(12)
    a = _kf_points(pts, "normal")
    taken, match = _state(frame, kp_taken)
    n = lib().orc_search_by_projection_sim3(frame._h, _ptr(scale), len(scale), logf(scale[1]), _ptr(cam), _ptr(tcw), len(a[0]),
                                            *[_ptr(x) for x in a], int(th), _ptr(taken), _ptr(match))
    return n, match[:frame.n]


# ----------------------------------------------------------------------------- front-end neighbours (numpy restatements)
def gray_from_color(img, rgb_order=True):
    """cvtColor(..., CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) of Tracking::GrabImage* (Tracking.cc:202-227),
    OpenCV's 8-bit path: (R*9798 + G*19235 + B*3735 + 16384) >> 15 (pinned to cv2 4.13 on all 2^24 colours in the tests)."""
    img = np.asarray(img, np.uint8)
    r, g, b = (img[..., 0], img[..., 1], img[..., 2]) if rgb_order else (img[..., 2], img[..., 1], img[..., 0])
    return ((r.astype(np.int64) * 9798 + g.astype(np.int64) * 19235 + b.astype(np.int64) * 3735 + 16384) >> 15).astype(np.uint8)


def depth_to_float(depth_u16, factor):
    """imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor), Tracking.cc:262-263: binary32 product of the value and the factor."""
    return (np.asarray(depth_u16).astype(np.float32) * np.float32(factor)).astype(np.float32)


def stereo_from_rgbd(keys, depth_f32, mbf):
    """Frame::ComputeStereoFromRGBD, Frame.cc:883-904 (mvKeysUn == mvKeys: no distortion)."""
    n = len(keys)
    ur = np.full(n, -1, np.float32)
    dp = np.full(n, -1, np.float32)
    for i in range(n):
        u, v = keys["x"][i], keys["y"][i]
        d = depth_f32[int(v), int(u)]                   # cv::Mat::at<float>(float, float): the indices truncate
        if d > 0:
            dp[i] = d
            ur[i] = np.float32(u - np.float32(np.float32(mbf) / d))
    return ur, dp


def _bow_args(side, with_ur):
    """ctypes arguments of one BowSide (keeps the converted arrays alive in the returned list)."""
    hdr = np.array([side["n"], len(side["node_id"])], np.int32)
    arrs = [hdr, np.ascontiguousarray(side["descriptors"], np.uint8), np.ascontiguousarray(side["keys_un"]),
            None if side.get("valid") is None else np.ascontiguousarray(side["valid"], np.uint8)]
    if with_ur:
        arrs.append(None if side.get("u_right") is None else np.ascontiguousarray(side["u_right"], np.float32))
    arrs += [np.ascontiguousarray(side["node_id"], np.uint32), np.ascontiguousarray(side["node_start"], np.int32),
             np.ascontiguousarray(side["node_idx"], np.int32)]
    return arrs, [None if a is None else _ptr(a) for a in arrs]


def search_by_bow(side1, side2, th_low=50, strict=False, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW, src/ORBmatcher.cc:159-288 (strict=False, read match21) and :522-655 (strict=True, read match12)."""
    k1, a1 = _bow_args(side1, False)
    k2, a2 = _bow_args(side2, False)
    m12 = np.empty(max(side1["n"], 1), np.int32); m21 = np.empty(max(side2["n"], 1), np.int32)
    n = lib().orc_search_by_bow(*a1, *a2, int(th_low), int(strict), nnratio, int(check_ori), _ptr(m12), _ptr(m21))
    return n, m12[:side1["n"]], m21[:side2["n"]]


def search_for_triangulation(side1, side2, f12, epipole, level_sigma2, scale_factors, only_stereo=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:657-823; valid = keypoint has no map point yet."""
    k1, a1 = _bow_args(side1, True)
    k2, a2 = _bow_args(side2, True)
    f = np.ascontiguousarray(f12, np.float32).reshape(9)
    s2 = np.ascontiguousarray(level_sigma2, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    m12 = np.empty(max(side1["n"], 1), np.int32)
    n = lib().orc_search_for_triangulation(*a1, *a2, _ptr(f), float(epipole[0]), float(epipole[1]), _ptr(s2), _ptr(sf),
                                           int(only_stereo), int(check_ori), _ptr(m12))
    return n, m12[:side1["n"]]


def distinctive_descriptors(desc, start):
    """MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:345-410, for a CSR of descriptor lists."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    start = np.ascontiguousarray(start, np.int32)
    best = np.empty(max(len(start) - 1, 1), np.int32)
    lib().orc_distinctive_descriptors(_ptr(desc), _ptr(start), len(start) - 1, _ptr(best))
    return best[:len(start) - 1]


def fuse_search(frame, scale, cam, tcw, pts, th, camera_centre=None, sim3=False):
    """Search half of ORBmatcher::Fuse (src/ORBmatcher.cc:825-966, sim3=False; :974-1100, sim3=True).  pts as for
    search_by_projection_sim3 (valid = pMP && !isBad() && !IsInKeyFrame).  Returns (best_idx, best_dist) per point."""
    scale = np.ascontiguousarray(scale, np.float32)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)          # Frame.cc / KeyFrame: mvInvLevelSigma2
    cam = np.ascontiguousarray(cam, np.float32)
    tcw = np.ascontiguousarray(tcw, np.float32).reshape(12)
    ow = minus_rt_t(tcw) if camera_centre is None else np.ascontiguousarray(camera_centre, np.float32)
    a = _kf_points(pts, "normal")
    n = len(a[0])
    bi = np.empty(max(n, 1), np.int32); bd = np.empty(max(n, 1), np.int32)
    lib().orc_fuse_search(frame._h, _ptr(scale), _ptr(inv_sigma2), len(scale), logf(scale[1]), _ptr(cam), _ptr(tcw), _ptr(ow), int(sim3), n,
                          *[_ptr(x) for x in a], float(th), _ptr(bi), _ptr(bd))
    return bi[:n], bd[:n]


def search_by_sim3(frame1, frame2, scale, cam, t1w, t2w, t21, t12, pts1, pts2, th):
    """ORBmatcher::SearchBySim3, src/ORBmatcher.cc:1102-1326.  pts1 / pts2: map points of the keyframes' keypoints (valid = present,
    not bad, not already matched); t21 = [sR21 | t21], t12 = [sR12 | t12] (12 floats each).  Returns (nFound, match12)."""
    scale = np.ascontiguousarray(scale, np.float32)
    cam = np.ascontiguousarray(cam, np.float32)
    T = [np.ascontiguousarray(t, np.float32).reshape(12) for t in (t1w, t2w, t21, t12)]
    def side(p):
        return [np.ascontiguousarray(p[k], t) for k, t in (("valid", np.uint8), ("world_pos", np.float32), ("min_distance", np.float32),
                                                           ("max_distance", np.float32), ("max_distance_raw", np.float32), ("descriptors", np.uint8))]
    a, b = side(pts1), side(pts2)
    m12 = np.empty(max(len(a[0]), 1), np.int32)
    n = lib().orc_search_by_sim3(frame1._h, frame2._h, _ptr(scale), len(scale), logf(scale[1]), _ptr(cam), *[_ptr(t) for t in T],
                                 len(a[0]), *[_ptr(x) for x in a], len(b[0]), *[_ptr(x) for x in b], float(th), _ptr(m12))
    return n, m12[:len(a[0])]


def assign_keypoints_to_masks(keys_un, depth, masks, th_depth, min_keypoints=5):
    """Head of Frame::BuildObject2DsRGBD / BuildObject2DsStereo (src/Frame.cc:240-385).  masks: [n_masks, h, w] uint8.
    Returns (mask_of_kp, mvObjectKpIndices [n, 2], object_of_mask, N_O)."""
    k = np.ascontiguousarray(keys_un); d = np.ascontiguousarray(depth, np.float32); mk = np.ascontiguousarray(masks, np.uint8)
    n, (nm, h, w) = len(k), mk.shape
    mo = np.empty(max(n, 1), np.int32); ok = np.empty((max(n, 1), 2), np.int32); om = np.empty(max(nm, 1), np.int32)
    no = lib().orc_assign_keypoints_to_masks(_ptr(k), _ptr(d), n, _ptr(mk), nm, w, h, float(th_depth), int(min_keypoints), _ptr(mo), _ptr(ok), _ptr(om))
    return mo[:n], ok[:n], om[:nm], no


def hsv_from_bgr(bgr):
    """cv::cvtColor(CV_BGR2HSV) on 8-bit pixels; bgr: [..., 3] uint8."""
    a = np.ascontiguousarray(bgr, np.uint8)
    out = np.empty_like(a)
    lib().orc_hsv_from_bgr(_ptr(a), a.size // 3, _ptr(out))
    return out


def hsv_histograms(im_bgr, masks):
    """Frame::ExtractHSVHistogramsFromMask (src/Frame.cc:388-414) per mask: [n_masks, 94] float32, V | S | H."""
    im = np.ascontiguousarray(im_bgr, np.uint8); mk = np.ascontiguousarray(masks, np.uint8)
    nm, h, w = mk.shape
    out = np.empty((nm, 94), np.float32)
    for m in range(nm):
        lib().orc_hsv_histogram(_ptr(im), _ptr(mk[m]), w, h, out[m].ctypes.data_as(C.c_void_p))
    return out


def undistort_points(pts, K, dist_coef):
    """cv::undistortPoints(pts, pts, K, distCoef, Mat(), K) as Frame::UndistortKeyPoints (src/Frame.cc:644-674) calls it; K = (fx, fy, cx, cy)."""
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); d = np.ascontiguousarray(dist_coef, np.float32)
    out = np.empty_like(p)
    lib().orc_undistort_points(_ptr(p), len(p), *[float(np.float32(v)) for v in K], _ptr(d), len(d), _ptr(out))
    return out


def distance_transform(masks):
    """cv::distanceTransform(~mask, DIST_L2, DIST_MASK_PRECISE) (src/ObjectTypes.cc:23) per mask: [n_masks, h, w] float32."""
    mk = np.ascontiguousarray(masks, np.uint8)
    nm, h, w = mk.shape
    out = np.empty((nm, h, w), np.float32)
    for m in range(nm):
        lib().orc_distance_transform(_ptr(mk[m]), w, h, out[m].ctypes.data_as(C.c_void_p))
    return out
