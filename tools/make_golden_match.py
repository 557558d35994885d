"""Writes tests/golden/match_*.npz: outputs of the matcher oracle (oracle/match_oracle.cpp) on seeded synthetic
cases.  The reference ships no fixtures for these functions and cannot be run here, so these are regression
fixtures of the restatement, not upstream golden vectors.  Usage: python tools/make_golden_match.py"""
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


def compute(kind, seed):
    shape = synth.TUM_SHAPE
    if kind == "map":
        frame, mp, kp_obs = mc.map_case(shape, 1000, 20000 if seed == 0 else 3000, seed, 0.0 if seed == 0 else 0.25)
        n, match = mc.oracle_map(frame, shape, mp, 3.0, 0.8, kp_obs)
        return dict(n_matches=np.int32(n), kp_match=match)
    if kind in ("last", "last_forward"):
        last, cur = synth.motion_pair(shape, 1000, seed, forward=0.6 if kind == "last_forward" else 0.0)
        n, match = mc.oracle_last(cur, shape, last, 7.0, False)
        return dict(n_matches=np.int32(n), kp_match=match)
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
            n, match = oracle.search_by_projection_sim3(F, synth.scale_factors(), synth.camera_for(shape), last["tcw_current"], pts, 10)
        return dict(n_matches=np.int32(n), kp_match=match)
    if kind == "knn2":
        D = synth.keyframe_descriptors(3, 2000, seed)
        bi, bd, sd = oracle.hamming_knn2(D[1], D[0], 50, 0.6)
        return dict(best_idx=bi, best_dist=bd, second_dist=sd)
    raise ValueError(kind)


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    for kind, seed in CASES:
        res = compute(kind, seed)
        np.savez_compressed(os.path.join(out, f"match_{kind}_s{seed}.npz"), kind=kind, seed=seed, **res)
        print(kind, seed, {k: (v.shape, int(np.sum(v >= 0)) if v.dtype.kind == "i" else None) for k, v in res.items()})
