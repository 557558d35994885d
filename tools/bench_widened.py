#!/usr/bin/env python
"""Throughput of the SURVEY 8(f) operators through the Python mirror (host arrays in and out, so every number includes the PCIe
copies and the packing of the host arrays), with the oracle restatement timed beside it on one host thread.  One JSON object.
usage (GPU box): python tools/bench_widened.py > gpurun_out/bench_widened.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from object_slam_b200 import synth
from object_slam_b200.matcher import ORBmatcher
from test_gpu_matchers import frame_set, oracle_frame


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    M = ORBmatcher(0.8, True)
    shape = synth.TUM_SHAPE
    out = {}
    # --- Fuse search / SearchBySim3: 16 keyframes x 1000 points
    B = 16
    pairs = [synth.motion_pair(shape, 1000, 70 + s) for s in range(B)]
    pts = [synth.keyframe_points(p[0], 75 + i) for i, p in enumerate(pairs)]
    fs = frame_set(M, shape, [p[1] for p in pairs])
    st = {k: np.stack([q[k] for q in pts]) for k in pts[0]}
    tcw = np.stack([p[0]["tcw_current"] for p in pairs]); ow = np.stack([oracle.minus_rt_t(t) for t in tcw])
    g = timed(lambda: M.FuseSearch(fs, st, tcw, 3.0, camera_centre=ow, per_frame=True), 20)
    c = timed(lambda: oracle.fuse_search(oracle_frame(pairs[0][1], shape), synth.scale_factors(), synth.camera_for(shape), tcw[0], pts[0], 3.0, camera_centre=ow[0]), 5)
    out["fuse_search"] = {"unit": "map points/s", "gpu": B * 1000 / g, "cpu_1_thread": 1000 / c, "batch": "16 keyframes x 1000 points"}
    cases = [synth.sim3_pair(shape, 1000, 40 + s) for s in range(B)]
    s1 = frame_set(M, shape, [c_[0] for c_ in cases]); s2 = frame_set(M, shape, [c_[1] for c_ in cases])
    stk = lambda i: {k: np.stack([c_[i][k] for c_ in cases]) for k in cases[0][i]}
    T = {k: np.stack([c_[4][k] for c_ in cases]) for k in cases[0][4]}
    p1, p2 = stk(2), stk(3)
    g = timed(lambda: M.SearchBySim3(s1, s2, p1, p2, T["t1w"], T["t2w"], T["t21"], T["t12"], 7.5, per_frame=True), 20)
    k1, k2, q1, q2, P = cases[0]
    c = timed(lambda: oracle.search_by_sim3(oracle_frame(k1, shape), oracle_frame(k2, shape), synth.scale_factors(), synth.camera_for(shape),
                                            P["t1w"], P["t2w"], P["t21"], P["t12"], q1, q2, 7.5), 5)
    out["search_by_sim3"] = {"unit": "keyframe pairs/s", "gpu": B / g, "cpu_1_thread": 1 / c, "batch": "16 pairs x 1000 keypoints"}
    # --- triangulation: 64 pairs x 2000 keypoints
    bp = [synth.bow_pair(synth.KITTI_SHAPE, 2000, 200 + i, n_nodes=100) for i in range(8)]
    A = [bp[i % 8][0] for i in range(64)]; Bs = [bp[i % 8][1] for i in range(64)]
    x = bp[0][2]
    f12 = np.stack([x["f12"]] * 64); ep = np.stack([np.array(x["epipole"], np.float32)] * 64)
    g = timed(lambda: M.SearchForTriangulation(A, Bs, f12, ep, x["level_sigma2"], x["scale_factors"]), 5)
    c = timed(lambda: oracle.search_for_triangulation(A[0], Bs[0], x["f12"], x["epipole"], x["level_sigma2"], x["scale_factors"], False, True), 5)
    # the same call with the keyframes packed once (what a C++ caller hands over): host arrays in, match12 out
    import ctypes as C
    from object_slam_b200._capi import check, lib, ptr
    sd1, keep1 = M._pack_side(A); sd2, keep2 = M._pack_side(Bs)
    s2l = np.ascontiguousarray(x["level_sigma2"], np.float32); sfl = np.ascontiguousarray(x["scale_factors"], np.float32)
    m12 = np.empty((64, sd1.cap), np.int32); nm = np.empty(64, np.int32)
    gp = timed(lambda: check(lib().obs_search_for_triangulation(M._h, C.byref(sd1), C.byref(sd2), 64, ptr(f12), ptr(ep), ptr(s2l), ptr(sfl), len(sfl), 0, 1,
                                                                ptr(m12), ptr(nm))), 10)
    out["search_for_triangulation_prepacked"] = {"unit": "keyframe pairs/s", "gpu": 64 / gp, "batch": "64 pairs x 2000 keypoints, arrays packed once; host arrays in, H2D inside the call"}
    out["search_for_triangulation"] = {"unit": "keyframe pairs/s", "gpu": 64 / g, "cpu_1_thread": 1 / c, "batch": "64 pairs x 2000 keypoints, 100 nodes",
                                       "note": "gpu time is dominated by packing 64 x 2 keyframes on the host (python) and the copies"}
    # --- distinctive descriptors: 20k map points
    dd, ds = synth.observation_descriptors(20000, 2, max_obs=24)
    g = timed(lambda: M.ComputeDistinctiveDescriptors(dd, ds), 10)
    c = timed(lambda: oracle.distinctive_descriptors(dd[:ds[2000]], ds[:2001]), 3)
    out["distinctive_descriptors"] = {"unit": "map points/s", "gpu": 20000 / g, "cpu_1_thread": 2000 / c, "batch": "20000 points, 1..24 observations"}
    # --- object layer on one TUM frame with 10 masks
    keys, _, _ = synth.synthetic_frame(shape, 1000, 3)
    depth = np.random.default_rng(3).uniform(0.2, 6, 1000).astype(np.float32)
    masks = synth.semantic_masks(shape, 10, 4)
    img = np.stack([synth.blocky_image(shape, 1), synth.blocky_image(shape, 2), synth.blocky_image(shape, 3)], -1)
    for name, gf, cf in (("assign_keypoints_to_masks", lambda: M.AssignKeypointsToMasks(keys, depth, masks, 3.5, 5), lambda: oracle.assign_keypoints_to_masks(keys, depth, masks, 3.5, 5)),
                         ("hsv_histograms", lambda: M.ExtractHSVHistogramsFromMasks(img, masks), lambda: oracle.hsv_histograms(img, masks)),
                         ("distance_transform", lambda: M.DistanceTransform(masks), lambda: oracle.distance_transform(masks)),
                         ("undistort_keypoints", lambda: M.UndistortKeyPoints(keys, (517.3, 516.5, 318.6, 255.3), np.array([0.26, -0.95, -0.005, 0.0026, 1.16], np.float32)),
                          lambda: oracle.undistort_points(np.stack([keys["x"], keys["y"]], 1), (517.3, 516.5, 318.6, 255.3), np.array([0.26, -0.95, -0.005, 0.0026, 1.16], np.float32)))):
        g = timed(gf, 20); c = timed(cf, 5)
        out[name] = {"unit": "frames/s (640x480, 1000 keypoints, 10 masks)", "gpu": 1 / g, "cpu_1_thread": 1 / c}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
