"""TEST INFRASTRUCTURE ONLY -- the reference's own matcher code as a checker.

``oracle/_ref/libref_matcher.so`` holds the reference's UNMODIFIED ``src/ORBmatcher.cc``, ``src/Frame.cc``, ``src/MapPoint.cc`` and
``src/KeyFrame.cc``, compiled in place from /root/reference against ``oracle/refstubs`` (stand-ins for the OpenCV / DBoW2 / g2o / Eigen
headers those files include) with ``oracle/ref_matcher_harness.cpp`` around them (oracle/Makefile).  The functions below take the
same flat arguments as the restatements in ``oracle/__init__.py`` (match_oracle.cpp), build the reference's own Frame / KeyFrame /
MapPoint objects from them, run the reference's methods and flatten what those wrote -- so a test is
``oracle.X(args) == oracle.refm.X(args)``, and the golden fixtures under tests/golden are generated from here.

The library only exists where /root/reference does (this container); the built file travels to the GPU box.
"""
import ctypes as C
import os

import numpy as np

from . import KEYPOINT_DTYPE, HERE, _ptr, logf

REFM_PATH = os.path.join(HERE, "_ref", "libref_matcher.so")


class RefFrame(C.Structure):
    _fields_ = [("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("n", C.c_int32),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float), ("mb", C.c_float),
                ("scale", C.c_void_p), ("nlevels", C.c_int32), ("log_scale_factor", C.c_float), ("tcw", C.c_void_p),
                ("n_nodes", C.c_int32), ("node_id", C.c_void_p), ("node_start", C.c_void_p), ("node_idx", C.c_void_p)]


class RefPoints(C.Structure):
    _fields_ = [("n", C.c_int32), ("valid", C.c_void_p), ("pos", C.c_void_p), ("min_dist_raw", C.c_void_p), ("max_dist_raw", C.c_void_p),
                ("normal", C.c_void_p), ("desc", C.c_void_p), ("obs", C.c_void_p)]


def available():
    return os.path.exists(REFM_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{REFM_PATH} is missing (it is built from /root/reference by `make -C oracle ref`)")
        L = C.CDLL(REFM_PATH)
        for name in ("refm_descriptor_distance", "refm_features_in_area", "refm_stereo_match", "refm_search_by_projection_map",
                     "refm_search_by_projection_last", "refm_search_for_initialization", "refm_search_by_projection_keyframe",
                     "refm_search_by_projection_sim3", "refm_search_by_bow", "refm_search_for_triangulation", "refm_search_by_sim3",
                     "refm_predict_scale"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _a(x):
    return None if x is None else C.c_void_p(x.ctypes.data)


class _Keep(list):
    """Arrays a ctypes structure points into."""


def frame(keys_un, desc, u_right=None, bounds=None, cam=None, scale=None, tcw=None, feat_vec=None):
    """(RefFrame, keep-alive list).  cam = (fx, fy, cx, cy, mbf, mb); tcw = 12 floats [R | t]; feat_vec = (node_id, node_start, node_idx)."""
    keep = _Keep()
    k = np.ascontiguousarray(keys_un, KEYPOINT_DTYPE); keep.append(k)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); keep.append(d)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32); keep.append(ur)
    sc = np.ascontiguousarray(scale if scale is not None else _default_scale(), np.float32); keep.append(sc)
    cam = (500.0, 500.0, 320.0, 240.0, 40.0, 0.08) if cam is None else [float(c) for c in cam]
    t = None if tcw is None else np.ascontiguousarray(tcw, np.float32).reshape(12); keep.append(t)
    f = RefFrame()
    f.keys_un, f.desc, f.u_right, f.n = k.ctypes.data, d.ctypes.data, (None if ur is None else ur.ctypes.data), len(k)
    f.min_x, f.max_x, f.min_y, f.max_y = [float(b) for b in bounds]
    f.fx, f.fy, f.cx, f.cy, f.mbf, f.mb = cam
    f.scale, f.nlevels = sc.ctypes.data, len(sc)
    f.log_scale_factor = logf(sc[1]) if len(sc) > 1 else 0.0
    f.tcw = None if t is None else t.ctypes.data
    if feat_vec is not None:
        nid = np.ascontiguousarray(feat_vec[0], np.uint32); ns = np.ascontiguousarray(feat_vec[1], np.int32); ni = np.ascontiguousarray(feat_vec[2], np.int32)
        keep += [nid, ns, ni]
        f.n_nodes, f.node_id, f.node_start, f.node_idx = len(nid), nid.ctypes.data, ns.ctypes.data, (ni.ctypes.data if len(ni) else None)
    keep.append(f)
    return f, keep


def _default_scale():
    s = np.ones(8, np.float32)
    for i in range(1, 8):
        s[i] = np.float32(s[i - 1] * 1.2)
    return s


def canonical_points(pts):
    """A keyframe-points dict as the reference's MapPoint can hold it.  The reference keeps two members per point (mfMinDistance,
    mfMaxDistance); GetMinDistanceInvariance = 0.8f * mfMinDistance, GetMaxDistanceInvariance = 1.2f * mfMaxDistance and PredictScale
    reads mfMaxDistance (MapPoint.cc:475-519).  synth.keyframe_points / synth.sim3_pair generate exactly that (they carry
    min_distance_raw / max_distance_raw); this checks it and is the identity then.  Dicts without the raw members get them recovered
    from the limits, and the limits recomputed, for BOTH sides of a comparison."""
    p = dict(pts)
    if pts.get("min_distance_raw") is None:
        p["min_distance_raw"] = (np.ascontiguousarray(pts["min_distance"], np.float32) / np.float32(0.8)).astype(np.float32)
        p["max_distance_raw"] = (np.ascontiguousarray(pts["max_distance"], np.float32) / np.float32(1.2)).astype(np.float32)
    p["min_distance"] = (np.float32(0.8) * np.ascontiguousarray(p["min_distance_raw"], np.float32)).astype(np.float32)
    p["max_distance"] = (np.float32(1.2) * np.ascontiguousarray(p["max_distance_raw"], np.float32)).astype(np.float32)
    if pts.get("min_distance_raw") is not None:
        assert np.array_equal(p["min_distance"], pts["min_distance"]) and np.array_equal(p["max_distance"], pts["max_distance"])
    return p


def points(pts, normal_key="normal", obs=None):
    """(RefPoints, keep-alive list) from a keyframe-points dict (valid, world_pos, min_distance_raw, max_distance_raw, normal, descriptors)."""
    keep = _Keep()
    def arr(key, dt):
        if pts.get(key) is None:
            return None
        a = np.ascontiguousarray(pts[key], dt); keep.append(a); return a
    valid, pos = arr("valid", np.uint8), arr("world_pos", np.float32)
    rmin, rmax = arr("min_distance_raw", np.float32), arr("max_distance_raw", np.float32)
    nrm, desc = arr(normal_key, np.float32) if normal_key else None, arr("descriptors", np.uint8)
    ob = None if obs is None else np.ascontiguousarray(obs, np.int32); keep.append(ob)
    p = RefPoints()
    p.n = len(desc) if desc is not None else len(valid)
    p.valid, p.pos, p.min_dist_raw, p.max_dist_raw = [None if a is None else a.ctypes.data for a in (valid, pos, rmin, rmax)]
    p.normal, p.desc, p.obs = [None if a is None else a.ctypes.data for a in (nrm, desc, ob)]
    keep.append(p)
    return p, keep


# ------------------------------------------------------------------------------------------------ the reference's functions
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().refm_descriptor_distance(_ptr(a), _ptr(b))


def compute_three_maxima(sizes):
    sizes = np.ascontiguousarray(sizes, np.int32)
    ind = np.empty(3, np.int32)
    lib().refm_compute_three_maxima(_ptr(sizes), len(sizes), _ptr(ind))
    return tuple(int(i) for i in ind)


def frame_grid(fr):
    f, keep = fr
    start = np.empty(64 * 48 + 1, np.int32); idx = np.empty(max(f.n, 1), np.int32)
    lib().refm_frame_grid(C.byref(f), _ptr(start), _ptr(idx))
    return start, idx[:start[-1]]


def features_in_area(fr, x, y, r, min_level=-1, max_level=-1, keyframe=False):
    f, keep = fr
    out = np.empty(max(f.n, 1), np.int32)
    n = lib().refm_features_in_area(C.byref(f), C.c_float(x), C.c_float(y), C.c_float(r), int(min_level), int(max_level), int(keyframe), _ptr(out), len(out))
    return out[:n].copy()


def stereo_match(kpsL, descL, kpsR, descR, pyrL, pyrR, scale, mbf, maxD, bounds):
    """Frame::ComputeStereoMatches as compiled.  The reference has minD = 0 and maxD = mbf / mb; mb is chosen so that mbf / mb is the
    binary32 nearest to maxD -- the value actually used is returned third, for the restatement to be called with."""
    mb = np.float32(np.float32(mbf) / np.float32(maxD))
    fr = frame(kpsL, descL, None, bounds, (1.0, 1.0, 0.0, 0.0, float(mbf), float(mb)), scale)
    f, keep = fr
    kR = np.ascontiguousarray(kpsR, KEYPOINT_DTYPE); dR = np.ascontiguousarray(descR, np.uint8).reshape(-1, 32)
    pl = [np.ascontiguousarray(p, np.uint8) for p in pyrL]; pr = [np.ascontiguousarray(p, np.uint8) for p in pyrR]
    n = len(pl)
    arrL = (C.c_void_p * n)(*[p.ctypes.data for p in pl]); arrR = (C.c_void_p * n)(*[p.ctypes.data for p in pr])
    lw = np.array([p.shape[1] for p in pl], np.int32); lh = np.array([p.shape[0] for p in pl], np.int32)
    ur = np.empty(max(f.n, 1), np.float32); dp = np.empty(max(f.n, 1), np.float32)
    used = C.c_float()
    lib().refm_stereo_match(C.byref(f), _ptr(kR), _ptr(dR), len(kR), arrL, arrR, _ptr(lw), _ptr(lh), _ptr(ur), _ptr(dp), C.byref(used))
    return ur[:f.n], dp[:f.n], float(used.value)


def search_by_projection_map(fr, in_view, proj_x, proj_y, proj_xr, level, view_cos, mp_desc, mp_obs, th, nnratio, kp_obs=None):
    f, keep = fr
    iv = np.ascontiguousarray(in_view, np.uint8)
    fl = [np.ascontiguousarray(a, np.float32) for a in (proj_x, proj_y, proj_xr)]
    lv = np.ascontiguousarray(level, np.int32); vc = np.ascontiguousarray(view_cos, np.float32)
    md = np.ascontiguousarray(mp_desc, np.uint8); mo = np.ascontiguousarray(mp_obs, np.int32)
    ko = None if kp_obs is None else np.ascontiguousarray(kp_obs, np.int32)
    match = np.empty(max(f.n, 1), np.int32)
    n = lib().refm_search_by_projection_map(C.byref(f), len(iv), _ptr(iv), _ptr(fl[0]), _ptr(fl[1]), _ptr(fl[2]), _ptr(lv), _ptr(vc), _ptr(md),
                                            _ptr(mo), C.c_float(th), C.c_float(nnratio), None if ko is None else _ptr(ko), _ptr(match))
    return n, match[:f.n]


def search_by_projection_last(fr_cur, tcw_last, last_has_point, last_pos, last_octave, last_angle, mp_desc, mp_obs, th, mono,
                              check_ori=True, kp_obs=None):
    """fr_cur carries the current frame's pose (tcw) and camera."""
    f, keep = fr_cur
    tl = np.ascontiguousarray(tcw_last, np.float32).reshape(12)
    has = np.ascontiguousarray(last_has_point, np.uint8); pos = np.ascontiguousarray(last_pos, np.float32).reshape(-1, 3)
    octv = np.ascontiguousarray(last_octave, np.int32); ang = np.ascontiguousarray(last_angle, np.float32)
    md = np.ascontiguousarray(mp_desc, np.uint8); mo = np.ascontiguousarray(mp_obs, np.int32)
    ko = None if kp_obs is None else np.ascontiguousarray(kp_obs, np.int32)
    match = np.empty(max(f.n, 1), np.int32)
    n = lib().refm_search_by_projection_last(C.byref(f), _ptr(tl), len(has), _ptr(has), _ptr(pos), _ptr(octv), _ptr(ang), _ptr(md), _ptr(mo),
                                             C.c_float(th), int(mono), int(check_ori), None if ko is None else _ptr(ko), _ptr(match))
    return n, match[:f.n]


def search_for_initialization(fr1, fr2, prev_matched, window, nnratio, check_ori=True):
    f1, k1 = fr1; f2, k2 = fr2
    pm = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.empty(max(f1.n, 1), np.int32)
    n = lib().refm_search_for_initialization(C.byref(f1), C.byref(f2), _ptr(pm), _ptr(m12), int(window), C.c_float(nnratio), int(check_ori))
    return n, m12[:f1.n], pm


def search_by_projection_keyframe(fr_cur, pts, th, orb_dist, check_ori=True, kp_taken=None):
    f, keep = fr_cur
    p, kp = points(pts, None)
    ang = np.ascontiguousarray(pts["angle"], np.float32)
    kt = None if kp_taken is None else np.ascontiguousarray(kp_taken, np.int32)
    match = np.empty(max(f.n, 1), np.int32)
    n = lib().refm_search_by_projection_keyframe(C.byref(f), C.byref(p), _ptr(ang), C.c_float(th), int(orb_dist), int(check_ori),
                                                 None if kt is None else _ptr(kt), _ptr(match))
    return n, match[:f.n]


def decompose_scw(scw):
    """([Rcw | tcw] as 12 floats, Ow) of ORBmatcher.cc:299-303 for an [sR | t] given as 12 floats."""
    s = np.ascontiguousarray(scw, np.float32).reshape(12)
    rt = np.empty(12, np.float32); ow = np.empty(3, np.float32)
    lib().refm_decompose_scw(_ptr(s), _ptr(rt), _ptr(ow))
    return rt, ow


def search_by_projection_sim3(fr_kf, scw, pts, th, kp_taken=None):
    f, keep = fr_kf
    p, kp = points(pts)
    s = np.ascontiguousarray(scw, np.float32).reshape(12)
    kt = None if kp_taken is None else np.ascontiguousarray(kp_taken, np.int32)
    match = np.empty(max(f.n, 1), np.int32)
    n = lib().refm_search_by_projection_sim3(C.byref(f), _ptr(s), C.byref(p), int(th), None if kt is None else _ptr(kt), _ptr(match))
    return n, match[:f.n]


def _bow_frame(side, bounds, cam=None, scale=None, tcw=None):
    return frame(side["keys_un"][:side["n"]], np.asarray(side["descriptors"])[:side["n"]],
                 None if side.get("u_right") is None else np.asarray(side["u_right"])[:side["n"]], bounds, cam, scale, tcw,
                 (side["node_id"], side["node_start"], side["node_idx"]))


def search_by_bow(side1, side2, bounds, keyframe_pair=False, nnratio=0.7, check_ori=True):
    """Returns (n, match12, match21) like oracle.search_by_bow (valid = the keyframe's keypoint has a good map point)."""
    f1, k1 = _bow_frame(side1, bounds); f2, k2 = _bow_frame(side2, bounds)
    v1 = None if side1.get("valid") is None else np.ascontiguousarray(side1["valid"], np.uint8)
    v2 = None if side2.get("valid") is None else np.ascontiguousarray(side2["valid"], np.uint8)
    m12 = np.empty(max(f1.n, 1), np.int32); m21 = np.empty(max(f2.n, 1), np.int32)
    n = lib().refm_search_by_bow(C.byref(f1), None if v1 is None else _ptr(v1), C.byref(f2), None if v2 is None else _ptr(v2), int(keyframe_pair),
                                 C.c_float(nnratio), int(check_ori), _ptr(m12), _ptr(m21))
    return n, m12[:f1.n], m21[:f2.n]


def search_for_triangulation(side1, side2, f12, bounds, cam, scale, tcw1, tcw2, only_stereo=False, check_ori=True):
    """Returns (n, match12, epipole): the epipole is what the reference derives from the two poses (ORBmatcher.cc:664-671).
    side["valid"] = keypoint has no map point yet (as in oracle.search_for_triangulation)."""
    f1, k1 = _bow_frame(side1, bounds, cam, scale, tcw1); f2, k2 = _bow_frame(side2, bounds, cam, scale, tcw2)
    h1 = None if side1.get("valid") is None else (1 - np.ascontiguousarray(side1["valid"], np.uint8)).astype(np.uint8)
    h2 = None if side2.get("valid") is None else (1 - np.ascontiguousarray(side2["valid"], np.uint8)).astype(np.uint8)
    f = np.ascontiguousarray(f12, np.float32).reshape(9)
    m12 = np.empty(max(f1.n, 1), np.int32); ep = np.empty(2, np.float32)
    n = lib().refm_search_for_triangulation(C.byref(f1), None if h1 is None else _ptr(h1), C.byref(f2), None if h2 is None else _ptr(h2), _ptr(f),
                                            int(only_stereo), int(check_ori), _ptr(m12), _ptr(ep))
    return n, m12[:f1.n], ep


def fuse(fr_kf, pts, th, scw=None):
    """ORBmatcher::Fuse per point (see ref_matcher_harness.cpp): (best_idx per point, count of one call over all points)."""
    f, keep = fr_kf
    p, kp = points(pts, obs=pts.get("observations"))
    s = None if scw is None else np.ascontiguousarray(scw, np.float32).reshape(12)
    bi = np.empty(max(p.n, 1), np.int32); total = C.c_int()
    lib().refm_fuse(C.byref(f), None if s is None else _ptr(s), int(scw is not None), C.byref(p), C.c_float(th), _ptr(bi), C.byref(total))
    return bi[:p.n], total.value


def sim3_transforms(s12, r12, t12):
    r = np.ascontiguousarray(r12, np.float32).reshape(9); t = np.ascontiguousarray(t12, np.float32).reshape(3)
    t21 = np.empty(12, np.float32); t12o = np.empty(12, np.float32)
    lib().refm_sim3_transforms(C.c_float(s12), _ptr(r), _ptr(t), _ptr(t21), _ptr(t12o))
    return t21, t12o


def search_by_sim3(fr1, fr2, pts1, pts2, s12, r12, t12, th, already=None):
    f1, k1 = fr1; f2, k2 = fr2
    p1, kp1 = points(pts1, None); p2, kp2 = points(pts2, None)
    r = np.ascontiguousarray(r12, np.float32).reshape(9); t = np.ascontiguousarray(t12, np.float32).reshape(3)
    al = None if already is None else np.ascontiguousarray(already, np.int32)
    m12 = np.empty(max(f1.n, 1), np.int32)
    n = lib().refm_search_by_sim3(C.byref(f1), C.byref(f2), C.byref(p1), C.byref(p2), C.c_float(s12), _ptr(r), _ptr(t), C.c_float(th),
                                  None if al is None else _ptr(al), _ptr(m12))
    return n, m12[:f1.n]


def distinctive_descriptors(desc, start, bounds=(0.0, 640.0, 0.0, 480.0)):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
    like, keep = frame(np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8), None, bounds)
    best = np.empty(max(len(start) - 1, 1), np.int32)
    lib().refm_distinctive_descriptors(C.byref(like), _ptr(desc), _ptr(start), len(start) - 1, _ptr(best))
    return best[:len(start) - 1]


def predict_scale(max_dist_raw, current_dist, scale=None, bounds=(0.0, 640.0, 0.0, 480.0)):
    like, keep = frame(np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8), None, bounds, None, scale)
    return lib().refm_predict_scale(C.byref(like), C.c_float(max_dist_raw), C.c_float(current_dist))
