#!/usr/bin/env python
"""Randomised parity sweep of the matchers on the GPU box (random sizes, thresholds, ratios, vocabulary sizes, validity patterns)
against the oracle, bit for bit.  usage: python tools/parity_sweep_match.py [rounds] > gpurun_out/parity_sweep_match.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
import matcher_cases as mc
from object_slam_b200 import synth
from object_slam_b200.matcher import ORBmatcher
from test_gpu_matchers import frame_set, oracle_frame

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(777)
M = ORBmatcher(0.8, True)
res = {}


def note(name, ok, info):
    r = res.setdefault(name, {"cases": 0, "mismatches": []})
    r["cases"] += 1
    if not ok:
        r["mismatches"].append(info)


t0 = time.time()
for it in range(rounds):
    shape = [synth.TUM_SHAPE, synth.KITTI_SHAPE, (240, 320)][int(rng.integers(0, 3))]
    n = int(rng.integers(40, 2600)); seed = int(rng.integers(0, 1 << 30))
    ratio = float(rng.choice([0.6, 0.7, 0.8, 0.9, 0.95])); ori = bool(rng.integers(0, 2))
    M.mfNNratio, M.mbCheckOrientation = ratio, ori
    info = dict(shape=shape, n=n, seed=seed, ratio=ratio, ori=ori)
    # SearchByProjection (map points)
    n_mp = int(rng.integers(1, 6000)); th = float(rng.choice([1.0, 3.0, 5.0, 15.0])); locked = float(rng.choice([0.0, 0.3]))
    frame, mp, kp_obs = mc.map_case(shape, n, n_mp, seed, kp_locked_fraction=locked)
    fs = frame_set(M, shape, [frame])
    kobs = None
    if kp_obs is not None:
        kobs = np.zeros((1, fs.cap), np.int32); kobs[0, :n] = kp_obs
    g, gm = M.SearchByProjection(fs, *[mp[k] for k in mc.MP_KEYS], th=th, kp_observations=kobs)
    on, om = mc.oracle_map(frame, shape, mp, th, ratio, kp_obs)
    note("search_by_projection", g[0] == on and np.array_equal(gm[0, :n], om), dict(info, n_mp=n_mp, th=th))
    # last frame / keyframe / sim3 projection / fuse
    last, cur = synth.motion_pair(shape, n, seed + 1, forward=float(rng.choice([0.0, 0.6, -0.6])))
    fs = frame_set(M, shape, [cur])
    th2 = float(rng.choice([7.0, 15.0, 3.0])); mono = bool(rng.integers(0, 2))
    g, gm = M.SearchByProjectionLast(fs, *[last[k] for k in mc.LAST_KEYS], last["tcw_last"], last["tcw_current"], th2, mono)
    on, om = mc.oracle_last(cur, shape, last, th2, mono, ori)
    note("search_by_projection_last", g[0] == on and np.array_equal(gm[0, :n], om), dict(info, th=th2, mono=mono))
    pts = synth.keyframe_points(last, seed + 2)
    F = oracle_frame(cur, shape)
    g, gm = M.SearchByProjectionKeyFrame(fs, pts, last["tcw_current"], th2, 100)
    on, om = oracle.search_by_projection_keyframe(F, synth.scale_factors(), synth.camera_for(shape), last["tcw_current"], pts, th2, 100, ori)
    note("search_by_projection_keyframe", g[0] == on and np.array_equal(gm[0, :n], om), dict(info, th=th2))
    thi = int(rng.choice([4, 10]))
    g, gm = M.SearchByProjectionSim3(fs, pts, last["tcw_current"], thi)
    on, om = oracle.search_by_projection_sim3(F, synth.scale_factors(), synth.camera_for(shape), last["tcw_current"], pts, thi)
    note("search_by_projection_sim3", g[0] == on and np.array_equal(gm[0, :n], om), dict(info, th=thi))
    ow = oracle.minus_rt_t(last["tcw_current"])
    for sim3 in (False, True):
        bi, bd = M.FuseSearch(fs, pts, last["tcw_current"][None], th2, camera_centre=None if sim3 else ow[None], sim3=sim3)
        oi, od = oracle.fuse_search(F, synth.scale_factors(), synth.camera_for(shape), last["tcw_current"], pts, th2, camera_centre=ow, sim3=sim3)
        note("fuse_search", np.array_equal(bi[0], oi) and np.array_equal(bd[0], od), dict(info, th=th2, sim3=sim3))
    # SearchBySim3
    k1, k2, p1, p2, P = synth.sim3_pair(shape, n, seed + 3)
    s1 = frame_set(M, shape, [k1]); s2 = frame_set(M, shape, [k2])
    g, gm = M.SearchBySim3(s1, s2, p1, p2, P["t1w"][None], P["t2w"][None], P["t21"][None], P["t12"][None], th2)
    on, om = oracle.search_by_sim3(oracle_frame(k1, shape), oracle_frame(k2, shape), synth.scale_factors(), synth.camera_for(shape),
                                   P["t1w"], P["t2w"], P["t21"], P["t12"], p1, p2, th2)
    note("search_by_sim3", g[0] == on and np.array_equal(gm[0], om), dict(info, th=th2))
    # SearchForInitialization
    f1, f2, prev = synth.init_pair(shape, n, seed + 4)
    a1 = frame_set(M, shape, [f1]); a2 = frame_set(M, shape, [f2])
    pv = np.zeros((1, a1.cap, 2), np.float32); pv[0, :n] = prev
    win = int(rng.choice([50, 100, 200]))
    g, gm = M.SearchForInitialization(a1, a2, pv, win)
    on, om, opm = mc.oracle_init(f1, f2, shape, prev, win, ratio, ori)
    note("search_for_initialization", g[0] == on and np.array_equal(gm[0, :n], om) and np.array_equal(pv[0, :n], opm), dict(info, window=win))
    # DBoW2-gated
    nodes = int(rng.choice([1, 3, 20, 100, 400]))
    a, b, x = synth.bow_pair(shape, n, seed + 5, n_nodes=nodes)
    for kf_pair in (False, True):
        g, m12, m21 = M.SearchByBoW([a], [b], keyframe_pair=kf_pair)
        on, o12, o21 = oracle.search_by_bow(a, b if kf_pair else dict(b, valid=None), 50, kf_pair, ratio, ori)
        note("search_by_bow", g[0] == on and np.array_equal(m12[0, :n], o12) and np.array_equal(m21[0, :n], o21), dict(info, nodes=nodes, kf_pair=kf_pair))
    only = bool(rng.integers(0, 2))
    g, m12 = M.SearchForTriangulation([a], [b], x["f12"], x["epipole"], x["level_sigma2"], x["scale_factors"], bOnlyStereo=only)
    on, o12 = oracle.search_for_triangulation(a, b, x["f12"], x["epipole"], x["level_sigma2"], x["scale_factors"], only, ori)
    note("search_for_triangulation", g[0] == on and np.array_equal(m12[0, :n], o12), dict(info, nodes=nodes, only_stereo=only))
    dd, ds = synth.observation_descriptors(int(rng.integers(1, 3000)), seed + 6, max_obs=int(rng.choice([3, 24, 70])))
    note("distinctive_descriptors", np.array_equal(M.ComputeDistinctiveDescriptors(dd, ds), oracle.distinctive_descriptors(dd, ds)), info)
    # brute force
    D = synth.keyframe_descriptors(3, n, seed + 7)
    M.mfNNratio = 0.6
    bi, bd, sd = M.knn2(D, np.array([(1, 0), (2, 1)], np.int32))
    ok = True
    for p, (q, dbi) in enumerate(((1, 0), (2, 1))):
        obi, obd, osd = oracle.hamming_knn2(D[q], D[dbi], 50, 0.6)
        ok &= np.array_equal(bi[p], obi) and np.array_equal(bd[p], obd) and np.array_equal(sd[p], osd)
    note("hamming_knn2", bool(ok), info)
res["_seconds"] = round(time.time() - t0, 1)
print(json.dumps(res, indent=1, default=str))
