"""Seeded matcher cases shared by the CPU oracle tests, the GPU parity tests and tools/make_golden_match.py."""
import numpy as np

import oracle
from object_slam_b200 import synth

MP_KEYS = ("in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "descriptors", "observations")
LAST_KEYS = ("has_point", "world_pos", "octave", "angle", "descriptors", "observations")


def bounds(shape):
    return (0.0, float(shape[1]), 0.0, float(shape[0]))       # mnMinX, mnMaxX, mnMinY, mnMaxY (Frame.cc:699-702)


def oracle_frame(frame, shape):
    return oracle.OracleFrame(frame[0], frame[1], frame[2], bounds(shape))


def map_case(shape, n_kp, n_mp, seed, kp_locked_fraction=0.0, anchored=0.5):
    frame = synth.synthetic_frame(shape, n_kp, seed)
    mp = synth.map_points_for_frame(frame[0], frame[1], shape, n_mp, seed + 1000, anchored=anchored)
    kp_obs = None
    if kp_locked_fraction > 0:
        rng = np.random.default_rng(seed + 2000)
        kp_obs = (rng.random(n_kp) < kp_locked_fraction).astype(np.int32) * rng.integers(1, 4, n_kp).astype(np.int32)
    return frame, mp, kp_obs


def oracle_map(frame, shape, mp, th, nnratio, kp_obs=None):
    F = oracle_frame(frame, shape)
    n, match, _ = oracle.search_by_projection_map(F, synth.scale_factors(), *[mp[k] for k in MP_KEYS], th, nnratio, kp_obs)
    return n, match


def oracle_last(cur, shape, last, th, mono, check_ori=True, kp_obs=None):
    F = oracle_frame(cur, shape)
    n, match, _ = oracle.search_by_projection_last(F, synth.scale_factors(), synth.camera_for(shape), last["tcw_current"],
                                                   last["tcw_last"], *[last[k] for k in LAST_KEYS], th, mono, check_ori, kp_obs)
    return n, match


def oracle_init(f1, f2, shape, prev, window, nnratio, check_ori=True):
    return oracle.search_for_initialization(oracle_frame(f1, shape), oracle_frame(f2, shape), prev, window, nnratio, check_ori)
