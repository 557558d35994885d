"""Writes tests/golden/match_*.npz from the REFERENCE ITSELF: the reference's own, unmodified src/ORBmatcher.cc / Frame.cc /
MapPoint.cc / KeyFrame.cc compiled in place (oracle/_ref/libref_matcher.so, oracle/refm.py) run on seeded synthetic cases; every
case is also computed by the restatement (oracle/match_oracle.cpp) and the two must agree before a fixture is written.  The
reference ships no fixtures of its own for these functions.  (knn2 is the candidate loop of SearchByBoW over whole keyframes, a
workload of this repository: its fixture comes from the restatement and numpy.)  Needs /root/reference; the fixtures travel.
Usage: python tools/make_golden_match.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from object_slam_b200 import synth  # noqa: E402
import matcher_cases as mc  # noqa: E402

CASES = [("map", 0), ("map", 1), ("last", 2), ("last_forward", 3), ("init", 4), ("knn2", 5), ("keyframe", 6), ("sim3", 7)]


def compute_reference(kind, seed):
    """The same cases through the compiled reference (oracle.refm); None for kinds the reference has no function for."""
    from oracle import refm
    shape = synth.TUM_SHAPE
    sf, cam = synth.scale_factors(), synth.camera_for(shape)
    rf = lambda fr, sh=shape, tcw=None: refm.frame(fr[0], fr[1], fr[2], mc.bounds(sh), synth.camera_for(sh), sf, tcw)
    if kind == "map":
        frame, mp, kp_obs = mc.map_case(shape, 1000, 20000 if seed == 0 else 3000, seed, 0.0 if seed == 0 else 0.25)
        n, match = refm.search_by_projection_map(rf(frame), *[mp[k] for k in mc.MP_KEYS], 3.0, 0.8, kp_obs)
        return dict(n_matches=np.int32(n), kp_match=match)
    if kind in ("last", "last_forward"):
        last, cur = synth.motion_pair(shape, 1000, seed, forward=0.6 if kind == "last_forward" else 0.0)
        n, match = refm.search_by_projection_last(rf(cur, tcw=last["tcw_current"]), last["tcw_last"], *[last[k] for k in mc.LAST_KEYS], 7.0, False)
        return dict(n_matches=np.int32(n), kp_match=match)
    if kind == "init":
        f1, f2, prev = synth.init_pair(synth.KITTI_SHAPE, 2000, seed)
        n, m12, pm = refm.search_for_initialization(rf(f1, synth.KITTI_SHAPE), rf(f2, synth.KITTI_SHAPE), prev, 100, 0.9)
        return dict(n_matches=np.int32(n), matches12=m12, prev_matched=pm)
    if kind in ("keyframe", "sim3"):
        last, cur = synth.motion_pair(shape, 1000, seed)
        pts = refm.canonical_points(synth.keyframe_points(last, seed + 100))
        if kind == "keyframe":
            n, match = refm.search_by_projection_keyframe(rf(cur, tcw=last["tcw_current"]), pts, 10.0, 100)
        else:
            n, match = refm.search_by_projection_sim3(rf(cur), last["tcw_current"], pts, 10)
            return dict(n_matches=np.int32(n), kp_match=match, rt_decomposed=refm.decompose_scw(last["tcw_current"])[0])
        return dict(n_matches=np.int32(n), kp_match=match)
    return None


def compute(kind, seed, rt_decomposed=None):
    """The cases through the restatement.  rt_decomposed: [Rcw | tcw] of the sim3 case as the reference decomposes Scw (stored in
    the fixture, so that machines without the compiled reference can replay it)."""
    shape = synth.TUM_SHAPE
    if kind == "map":
        frame, mp, kp_obs = mc.map_case(shape, 1000, 20000 if seed == 0 else 3000, seed, 0.0 if seed == 0 else 0.25)
        n, match = mc.oracle_map(frame, shape, mp, 3.0, 0.8, kp_obs)
        return dict(n_matches=np.int32(n), kp_match=match)
    if kind in ("last", "last_forward"):
        last, cur = synth.motion_pair(shape, 1000, seed, forward=0.6 if kind == "last_forward" else 0.0)
        n, match = mc.oracle_last(cur, shape, last, 7.0, False)
        return dict(n_matches=np.int32(n), kp_match=np.where(match == -2, -1, match))     # -2: removed by the rotation check (restatement only)
    if kind == "init":
        f1, f2, prev = synth.init_pair(synth.KITTI_SHAPE, 2000, seed)
        n, m12, pm = mc.oracle_init(f1, f2, synth.KITTI_SHAPE, prev, 100, 0.9)
        return dict(n_matches=np.int32(n), matches12=m12, prev_matched=pm)
    if kind in ("keyframe", "sim3"):
        last, cur = synth.motion_pair(shape, 1000, seed)
        pts = synth.keyframe_points(last, seed + 100)
        F = mc.oracle_frame(cur, shape)
        if kind == "keyframe":
            n, match = oracle.search_by_projection_keyframe(F, synth.scale_factors(), synth.camera_for(shape), last["tcw_current"], pts, 10.0, 100)
        else:
            # SearchByProjection(KeyFrame*, Scw, ...) decomposes Scw itself (ORBmatcher.cc:299-303); the restatement starts behind that
            if rt_decomposed is None:
                from oracle import refm
                rt_decomposed = refm.decompose_scw(last["tcw_current"])[0]
            n, match = oracle.search_by_projection_sim3(F, synth.scale_factors(), synth.camera_for(shape), rt_decomposed, pts, 10)
            return dict(n_matches=np.int32(n), kp_match=np.where(match == -2, -1, match), rt_decomposed=np.asarray(rt_decomposed, np.float32))
        return dict(n_matches=np.int32(n), kp_match=np.where(match == -2, -1, match))
    if kind == "knn2":
        D = synth.keyframe_descriptors(3, 2000, seed)
        bi, bd, sd = oracle.hamming_knn2(D[1], D[0], 50, 0.6)
        return dict(best_idx=bi, best_dist=bd, second_dist=sd)
    raise ValueError(kind)


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    from oracle import refm
    assert refm.available(), "needs oracle/_ref/libref_matcher.so (make -C oracle ref, with /root/reference present)"
    for kind, seed in CASES:
        res = compute(kind, seed)
        ref = compute_reference(kind, seed)
        if ref is not None:
            for k in ref:
                assert np.array_equal(ref[k], res[k]), (kind, seed, k, "restatement differs from the compiled reference")
            res = ref
        np.savez_compressed(os.path.join(out, f"match_{kind}_s{seed}.npz"), kind=kind, seed=seed, **res)
        print(kind, seed, {k: (v.shape, int(np.sum(v >= 0)) if v.dtype.kind == "i" else None) for k, v in res.items()})
