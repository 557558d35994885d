"""DBoW2-gated matchers (SearchByBoW x2, SearchForTriangulation) and ComputeDistinctiveDescriptors.

CPU: the C++ oracle restatement (oracle/match_oracle.cpp) against an independent pure-Python restatement that
walks std::map-like dicts the way the reference does, and against the committed fixtures.  GPU: the CUDA kernels
(csrc/bow.cu) through the C ABI against the oracle, bit-exact (integer indices and counts)."""
import os

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bow_cases.npz")
F32 = np.float32


def _popcount(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _rot_bin(a1, a2):
    rot = F32(a1) - F32(a2)
    if rot < 0.0:
        rot = F32(rot + F32(360.0))
    v = float(F32(rot * F32(1.0 / 30)))
    b = int(np.floor(v + 0.5)) if v >= 0 else int(np.ceil(v - 0.5))      # C round(): half away from zero
    return 0 if b == 30 else b


def _three_maxima(sizes):
    max1 = max2 = max3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            i3, i2, i1 = i2, i1, i
        elif s > max2:
            max3, max2 = max2, s
            i3, i2 = i2, i
        elif s > max3:
            max3, i3 = s, i
    if max2 < F32(0.1) * F32(max1):
        i2 = i3 = -1
    elif max3 < F32(0.1) * F32(max1):
        i3 = -1
    return i1, i2, i3


def _featvec(side):
    """std::map<NodeId, vector<unsigned>>"""
    return {int(side["node_id"][k]): [int(i) for i in side["node_idx"][side["node_start"][k]:side["node_start"][k + 1]]]
            for k in range(len(side["node_id"]))}


def _walk(fv1, fv2):
    """the reference's two-iterator walk with lower_bound jumps: yields the vectors of the common nodes, ascending"""
    k1, k2 = sorted(fv1), sorted(fv2)
    a = b = 0
    while a < len(k1) and b < len(k2):
        if k1[a] == k2[b]:
            yield fv1[k1[a]], fv2[k2[b]]
            a += 1; b += 1
        elif k1[a] < k2[b]:
            a = int(np.searchsorted(k1, k2[b], "left"))
        else:
            b = int(np.searchsorted(k2, k1[a], "left"))


def py_search_by_bow(s1, s2, th_low, strict, nnratio, check_ori):
    m12 = -np.ones(s1["n"], np.int32); m21 = -np.ones(s2["n"], np.int32)
    hist = [[] for _ in range(30)]
    n = 0
    for v1, v2 in _walk(_featvec(s1), _featvec(s2)):
        for i1 in v1:
            if s1.get("valid") is not None and not s1["valid"][i1]:
                continue
            b1, bi, b2 = 256, -1, 256
            for i2 in v2:
                if m21[i2] >= 0 or (s2.get("valid") is not None and not s2["valid"][i2]):
                    continue
                d = _popcount(s1["descriptors"][i1], s2["descriptors"][i2])
                if d < b1:
                    b2, b1, bi = b1, d, i2
                elif d < b2:
                    b2 = d
            if (b1 < th_low if strict else b1 <= th_low) and F32(b1) < F32(nnratio) * F32(b2):
                m12[i1] = bi; m21[bi] = i1; n += 1
                if check_ori:
                    hist[_rot_bin(s1["keys_un"]["angle"][i1], s2["keys_un"]["angle"][bi])].append(i1)
    if check_ori:
        keep = _three_maxima([len(h) for h in hist])
        for b, h in enumerate(hist):
            if b in keep:
                continue
            for i1 in h:
                m21[m12[i1]] = -1; m12[i1] = -1; n -= 1
    return n, m12, m21


def py_search_for_triangulation(s1, s2, f12, epi, sigma2, sf, only_stereo, check_ori):
    F = np.asarray(f12, F32).reshape(3, 3)
    ex, ey = F32(epi[0]), F32(epi[1])
    m12 = -np.ones(s1["n"], np.int32)
    hist = [[] for _ in range(30)]
    n = 0
    k1, k2 = s1["keys_un"], s2["keys_un"]
    for v1, v2 in _walk(_featvec(s1), _featvec(s2)):
        for i1 in v1:
            if not s1["valid"][i1]:
                continue
            st1 = s1["u_right"][i1] >= 0
            if only_stereo and not st1:
                continue
            x1, y1 = k1["x"][i1], k1["y"][i1]
            best, bi = 50, -1
            for i2 in v2:
                if not s2["valid"][i2]:
                    continue
                st2 = s2["u_right"][i2] >= 0
                if only_stereo and not st2:
                    continue
                d = _popcount(s1["descriptors"][i1], s2["descriptors"][i2])
                if d > 50 or d > best:
                    continue
                x2, y2, o2 = k2["x"][i2], k2["y"][i2], k2["octave"][i2]
                if not st1 and not st2:
                    dx, dy = F32(ex - x2), F32(ey - y2)
                    if F32(F32(dx * dx) + F32(dy * dy)) < F32(F32(100) * sf[o2]):
                        continue
                a = F32(F32(F32(x1 * F[0, 0]) + F32(y1 * F[1, 0])) + F[2, 0])
                b = F32(F32(F32(x1 * F[0, 1]) + F32(y1 * F[1, 1])) + F[2, 1])
                c = F32(F32(F32(x1 * F[0, 2]) + F32(y1 * F[1, 2])) + F[2, 2])
                num = F32(F32(F32(a * x2) + F32(b * y2)) + c)
                den = F32(F32(a * a) + F32(b * b))
                if den == 0:
                    continue
                dsqr = F32(F32(num * num) / den)
                if float(dsqr) < 3.84 * float(sigma2[o2]):
                    bi, best = i2, d
            if bi >= 0:
                m12[i1] = bi; n += 1
                if check_ori:
                    hist[_rot_bin(k1["angle"][i1], k2["angle"][bi])].append(i1)
    if check_ori:
        keep = _three_maxima([len(h) for h in hist])
        for b, h in enumerate(hist):
            if b not in keep:
                for i1 in h:
                    m12[i1] = -1; n -= 1
    return n, m12


def py_distinctive(desc, start):
    out = []
    for p in range(len(start) - 1):
        d = desc[start[p]:start[p + 1]]
        N = len(d)
        if N == 0:
            out.append(-1); continue
        D = np.array([[_popcount(d[i], d[j]) for j in range(N)] for i in range(N)])
        med = [sorted(D[i])[int(0.5 * (N - 1))] for i in range(N)]
        out.append(int(np.argmin(med)))          # first minimum
    return np.array(out, np.int32)


CASES = [(120, 3, 12), (300, 4, 40), (300, 5, 7)]        # keypoints, seed, vocabulary nodes


@pytest.mark.parametrize("n,seed,nodes", CASES)
def test_oracle_vs_python_bow(n, seed, nodes):
    a, b, x = synth.bow_pair(synth.TUM_SHAPE, n, seed, n_nodes=nodes)
    for strict, ratio, ori in ((False, 0.7, True), (True, 0.8, True), (True, 0.6, False)):
        b2 = b if strict else dict(b, valid=None)
        got = oracle.search_by_bow(a, b2, 50, strict, ratio, ori)
        want = py_search_by_bow(a, b2, 50, strict, ratio, ori)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        assert got[0] > 0


@pytest.mark.parametrize("n,seed,nodes", CASES)
def test_oracle_vs_python_triangulation(n, seed, nodes):
    a, b, x = synth.bow_pair(synth.TUM_SHAPE, n, seed, n_nodes=nodes)
    for only_stereo, ori in ((False, True), (True, True), (False, False)):
        got = oracle.search_for_triangulation(a, b, x["f12"], x["epipole"], x["level_sigma2"], x["scale_factors"], only_stereo, ori)
        want = py_search_for_triangulation(a, b, x["f12"], x["epipole"], x["level_sigma2"], x["scale_factors"], only_stereo, ori)
        assert got[0] == want[0] and np.array_equal(got[1], want[1])
    assert got[0] > 0


def test_oracle_vs_python_distinctive():
    d, s = synth.observation_descriptors(60, 9, max_obs=12)
    assert np.array_equal(oracle.distinctive_descriptors(d, s), py_distinctive(d, s))


def golden_poses(seed):
    """Keyframe poses of the triangulation fixture (the reference derives the epipole from them, ORBmatcher.cc:664-671)."""
    rng = np.random.default_rng(seed)
    t1 = np.hstack([synth._rot(*rng.normal(0, 0.02, 3)), rng.normal(0, 0.3, (3, 1))]).astype(np.float32)
    t2 = np.hstack([synth._rot(*rng.normal(0, 0.02, 3)), rng.normal(0, 0.3, (3, 1))]).astype(np.float32)
    return t1, t2


def golden_cases(reference=False, stored=None):
    """The fixture's cases through the restatement, or (reference=True, tools/make_golden_bow.py) through the reference's own
    ORBmatcher.cc / MapPoint.cc compiled in place (oracle.refm).  stored: a loaded fixture, whose epipoles replay the triangulation
    cases where the compiled reference is not available."""
    from matcher_cases import bounds
    out = {}
    shape = synth.TUM_SHAPE
    for n, seed, nodes in [(500, 11, 60), (1000, 12, 100)]:
        a, b, x = synth.bow_pair(shape, n, seed, n_nodes=nodes)
        k = f"{n}_{seed}"
        t1, t2 = golden_poses(seed)
        if reference:
            from oracle import refm
            r = refm.search_by_bow(a, dict(b, valid=None), bounds(shape), False, 0.7, True)
            out[f"bowA_n_{k}"] = np.int32(r[0]); out[f"bowA_m21_{k}"] = r[2]
            r = refm.search_by_bow(a, b, bounds(shape), True, 0.8, True)
            out[f"bowB_n_{k}"] = np.int32(r[0]); out[f"bowB_m12_{k}"] = r[1]
            r = refm.search_for_triangulation(a, b, x["f12"], bounds(shape), synth.camera_for(shape), x["scale_factors"], t1, t2, False, True)
            out[f"tri_n_{k}"] = np.int32(r[0]); out[f"tri_m12_{k}"] = r[1]; out[f"tri_epipole_{k}"] = r[2]
        else:
            r = oracle.search_by_bow(a, dict(b, valid=None), 50, False, 0.7, True)
            out[f"bowA_n_{k}"] = np.int32(r[0]); out[f"bowA_m21_{k}"] = r[2]
            r = oracle.search_by_bow(a, b, 50, True, 0.8, True)
            out[f"bowB_n_{k}"] = np.int32(r[0]); out[f"bowB_m12_{k}"] = r[1]
            ep = stored[f"tri_epipole_{k}"]
            r = oracle.search_for_triangulation(a, b, x["f12"], ep, x["level_sigma2"], x["scale_factors"], False, True)
            out[f"tri_n_{k}"] = np.int32(r[0]); out[f"tri_m12_{k}"] = r[1]; out[f"tri_epipole_{k}"] = np.asarray(ep, np.float32)
    d, s = synth.observation_descriptors(500, 13)
    dd = np.asarray(d).reshape(-1, 32)
    best = (oracle.refm.distinctive_descriptors(d, s) if reference else oracle.distinctive_descriptors(d, s))
    # the chosen descriptor itself (an observation list may hold equal descriptors under different indices)
    out["distinctive"] = np.stack([dd[s[p] + best[p]] if best[p] >= 0 else np.zeros(32, np.uint8) for p in range(len(s) - 1)])
    return out


def test_oracle_matches_golden():
    """fixtures written by tools/make_golden_bow.py from this oracle at the commit that introduced it"""
    want = np.load(GOLDEN)
    got = golden_cases(stored=want)
    assert sorted(want.files) == sorted(got)
    for k in want.files:
        assert np.array_equal(want[k], got[k]), k


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def matcher():
    from object_slam_b200.matcher import ORBmatcher
    m = ORBmatcher(0.7, True)
    yield m
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,nodes,B", [(300, 25, 3), (1000, 100, 4), (2000, 100, 2), (64, 1, 2)])
def test_gpu_search_by_bow(matcher, n, nodes, B):
    pairs = [synth.bow_pair(synth.TUM_SHAPE, n, 100 + i, n_nodes=nodes) for i in range(B)]
    for kf_pair, ratio, ori in ((False, 0.7, True), (True, 0.8, True), (False, 0.9, False)):
        matcher.mfNNratio, matcher.mbCheckOrientation = ratio, ori
        nm, m12, m21 = matcher.SearchByBoW([p[0] for p in pairs], [p[1] for p in pairs], keyframe_pair=kf_pair)
        for i, (a, b, _) in enumerate(pairs):
            on, o12, o21 = oracle.search_by_bow(a, b if kf_pair else dict(b, valid=None), 50, kf_pair, ratio, ori)
            assert nm[i] == on
            assert np.array_equal(m12[i, :n], o12) and np.array_equal(m21[i, :n], o21)
    matcher.mfNNratio, matcher.mbCheckOrientation = 0.7, True


@pytest.mark.gpu
@pytest.mark.parametrize("n,nodes,B", [(300, 25, 3), (1000, 100, 4), (2000, 100, 2)])
def test_gpu_search_for_triangulation(matcher, n, nodes, B):
    pairs = [synth.bow_pair(synth.TUM_SHAPE, n, 200 + i, n_nodes=nodes) for i in range(B)]
    x = pairs[0][2]
    f12 = np.stack([p[2]["f12"] for p in pairs]); ep = np.stack([np.array(p[2]["epipole"], np.float32) for p in pairs])
    for only_stereo, ori in ((False, True), (True, True), (False, False)):
        matcher.mbCheckOrientation = ori
        nm, m12 = matcher.SearchForTriangulation([p[0] for p in pairs], [p[1] for p in pairs], f12, ep, x["level_sigma2"],
                                                 x["scale_factors"], bOnlyStereo=only_stereo)
        for i, (a, b, e) in enumerate(pairs):
            on, o12 = oracle.search_for_triangulation(a, b, e["f12"], e["epipole"], e["level_sigma2"], e["scale_factors"], only_stereo, ori)
            assert nm[i] == on and np.array_equal(m12[i, :n], o12)
    matcher.mbCheckOrientation = True


@pytest.mark.gpu
def test_gpu_distinctive_descriptors(matcher):
    for n_points, seed, mx in ((500, 1, 24), (3000, 2, 40), (10, 3, 70)):
        d, s = synth.observation_descriptors(n_points, seed, max_obs=mx)
        assert np.array_equal(matcher.ComputeDistinctiveDescriptors(d, s), oracle.distinctive_descriptors(d, s))


@pytest.mark.gpu
def test_gpu_bow_matches_golden(matcher):
    want = np.load(GOLDEN)
    for n, seed, nodes in [(500, 11, 60), (1000, 12, 100)]:
        a, b, x = synth.bow_pair(synth.TUM_SHAPE, n, seed, n_nodes=nodes)
        k = f"{n}_{seed}"
        matcher.mfNNratio = 0.7
        nm, m12, m21 = matcher.SearchByBoW([a], [b], keyframe_pair=False)
        assert nm[0] == want[f"bowA_n_{k}"] and np.array_equal(m21[0, :n], want[f"bowA_m21_{k}"])
        matcher.mfNNratio = 0.8
        nm, m12, m21 = matcher.SearchByBoW([a], [b], keyframe_pair=True)
        assert nm[0] == want[f"bowB_n_{k}"] and np.array_equal(m12[0, :n], want[f"bowB_m12_{k}"])
        # the fixture's epipole is the one the reference derived from the fixture's poses (ORBmatcher.cc:664-671)
        nm, m12 = matcher.SearchForTriangulation([a], [b], x["f12"], want[f"tri_epipole_{k}"], x["level_sigma2"], x["scale_factors"])
        assert nm[0] == want[f"tri_n_{k}"] and np.array_equal(m12[0, :n], want[f"tri_m12_{k}"])
    matcher.mfNNratio = 0.7
    d, s = synth.observation_descriptors(500, 13)
    best = matcher.ComputeDistinctiveDescriptors(d, s)
    dd = np.asarray(d).reshape(-1, 32)
    got = np.stack([dd[s[p] + best[p]] if best[p] >= 0 else np.zeros(32, np.uint8) for p in range(len(s) - 1)])
    assert np.array_equal(got, want["distinctive"])


# ------------------------------------------------------------------------------------------------ object layer
def py_assign(keys, depth, masks, th_depth, min_kp):
    """independent restatement: per keypoint the first accepting mask, then counts / ranks"""
    n = len(keys); nm, h, w = masks.shape
    first = -np.ones(n, np.int32)
    for k in range(n):
        if not (depth[k] > 0 and depth[k] <= th_depth):
            continue
        iy = (np.float32(keys["y"][k]) + np.arange(-10, 10, dtype=np.float32)).astype(np.int64)
        ix = (np.float32(keys["x"][k]) + np.arange(-10, 10, dtype=np.float32)).astype(np.int64)
        if iy.min() < 0 or iy.max() >= h or ix.min() < 0 or ix.max() >= w:
            continue
        for m in range(nm):
            if (masks[m][np.ix_(iy, ix)] == 255).all():
                first[k] = m
                break
    okp = -np.ones((n, 2), np.int32); om = -np.ones(nm, np.int32); nobj = 0
    for m in range(nm):
        idx = np.flatnonzero(first == m)
        if len(idx) > min_kp:
            okp[idx, 0] = nobj; okp[idx, 1] = np.arange(len(idx)); om[m] = nobj; nobj += 1
    return first, okp, om, nobj


def _mask_case(seed, n=1500, n_masks=9):
    keys, _, _ = synth.synthetic_frame(synth.TUM_SHAPE, n, seed)
    rng = np.random.default_rng(seed)
    depth = rng.uniform(-0.5, 6.0, n).astype(np.float32)
    return keys, depth, synth.semantic_masks(synth.TUM_SHAPE, n_masks, seed + 1)


@pytest.mark.parametrize("seed,min_kp", [(1, 5), (2, 10), (3, 5)])
def test_oracle_vs_python_mask_assignment(seed, min_kp):
    keys, depth, masks = _mask_case(seed)
    got = oracle.assign_keypoints_to_masks(keys, depth, masks, 3.5, min_kp)
    want = py_assign(keys, depth, masks, 3.5, min_kp)
    assert all(np.array_equal(a, b) for a, b in zip(got[:3], want[:3])) and got[3] == want[3]
    assert got[3] >= 1 and (got[0] >= 0).sum() > 20


@pytest.mark.gpu
@pytest.mark.parametrize("seed,min_kp,n", [(1, 5, 1500), (2, 10, 1500), (4, 5, 2500), (5, 5, 40)])
def test_gpu_mask_assignment(matcher, seed, min_kp, n):
    keys, depth, masks = _mask_case(seed, n=n)
    got = matcher.AssignKeypointsToMasks(keys, depth, masks, 3.5, min_kp)
    want = oracle.assign_keypoints_to_masks(keys, depth, masks, 3.5, min_kp)
    assert all(np.array_equal(a, b) for a, b in zip(got[:3], want[:3])) and got[3] == want[3]


def _cv2():
    try:
        import cv2
        return cv2
    except Exception:
        pytest.skip("cv2 not importable")


def test_oracle_hsv_matches_cv2_on_all_colours():
    """the restated 8-bit BGR->HSV against cv2.cvtColor over all 2^24 colours"""
    cv2 = _cv2()
    g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for r in range(0, 256):
        img = np.stack([b, g, np.full_like(b, r)], -1)
        assert np.array_equal(oracle.hsv_from_bgr(img), cv2.cvtColor(img, cv2.COLOR_BGR2HSV)), r


def _colour_image(shape, seed):
    rng = np.random.default_rng(seed)
    base = synth.blocky_image(shape, seed)
    return np.stack([base, np.roll(base, 7, 1) // 2 + rng.integers(0, 90, shape).astype(np.uint8),
                     255 - np.roll(base, 13, 0)], -1).astype(np.uint8)


def test_oracle_hsv_histogram_matches_cv2():
    """ExtractHSVHistogramsFromMask restated vs the same cv2 calls (cvtColor, calcHist x3, hconcat order, normalize L1)"""
    cv2 = _cv2()
    img = _colour_image(synth.TUM_SHAPE, 3)
    masks = synth.semantic_masks(synth.TUM_SHAPE, 6, 8)
    masks[5] = 0                                               # an empty mask: all-zero histogram
    got = oracle.hsv_histograms(img, masks)
    hsv = cv2.cvtColor(img.copy(), cv2.COLOR_BGR2HSV)
    for m in range(len(masks)):
        H = None
        for ch, size, rng in ((0, 30, [0, 180]), (1, 32, [0, 256]), (2, 32, [0, 256])):
            hist = cv2.calcHist([hsv], [ch], masks[m], [size], rng, accumulate=False).reshape(1, -1)
            H = hist.copy() if H is None else cv2.hconcat([hist, H])
        H = cv2.normalize(H, None, norm_type=cv2.NORM_L1) if H.sum() > 0 else H
        assert np.array_equal(got[m], H.reshape(-1)), m


@pytest.mark.gpu
def test_gpu_hsv_histograms(matcher):
    for seed, shape in ((3, synth.TUM_SHAPE), (4, synth.KITTI_SHAPE)):
        img = _colour_image(shape, seed)
        masks = synth.semantic_masks(shape, 7, seed + 5)
        masks[6] = 0
        assert np.array_equal(matcher.ExtractHSVHistogramsFromMasks(img, masks), oracle.hsv_histograms(img, masks))


TUM1_K = (517.306408, 516.469215, 318.643040, 255.313989)                  # Examples/RGB-D/TUM1.yaml
DISTORTIONS = [np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32),     # TUM1.yaml k1 k2 p1 p2 k3
               np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)]    # EuRoC.yaml k1 k2 p1 p2


@pytest.mark.parametrize("dist", DISTORTIONS, ids=["tum1", "euroc"])
def test_oracle_undistort_matches_cv2(dist):
    cv2 = _cv2()
    rng = np.random.default_rng(1)
    pts = np.stack([rng.uniform(0, 640, 200000), rng.uniform(0, 480, 200000)], 1).astype(np.float32)
    fx, fy, cx, cy = TUM1_K
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
    assert np.array_equal(oracle.undistort_points(pts, TUM1_K, dist), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("dist", DISTORTIONS + [np.zeros(5, np.float32)], ids=["tum1", "euroc", "none"])
def test_gpu_undistort_keypoints(matcher, dist):
    keys, _, _ = synth.synthetic_frame(synth.TUM_SHAPE, 2000, 9)
    un = matcher.UndistortKeyPoints(keys, TUM1_K, dist)
    want = keys.copy()
    if dist[0] != 0:
        p = oracle.undistort_points(np.stack([keys["x"], keys["y"]], 1), TUM1_K, dist)
        want["x"], want["y"] = p[:, 0], p[:, 1]
    assert un.tobytes() == want.tobytes()
    corners = np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32)            # Frame::ComputeImageBounds
    assert np.array_equal(matcher.UndistortPoints(corners, TUM1_K, dist), oracle.undistort_points(corners, TUM1_K, dist))


def _dt_masks(shape, seed):
    masks = synth.semantic_masks(shape, 6, seed)
    masks = np.concatenate([masks, np.zeros((1,) + shape, np.uint8), np.full((1,) + shape, 255, np.uint8)])
    masks[2][masks[2] == 0] = 100                             # values other than 0 / 255 are background too (~100 != 0)
    masks[3, 5, 7] = 255                                       # an isolated object pixel
    return masks


@pytest.mark.parametrize("shape", [(90, 130), (61, 77), (120, 160)])
def test_oracle_distance_transform_matches_cv2(shape):
    """trueDistTrans restated vs cv2.distanceTransform with IPP off (IPP's variant differs in the last bit on some sizes)"""
    cv2 = _cv2()
    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        masks = _dt_masks(shape, 5)
        got = oracle.distance_transform(masks)
        for m in range(len(masks)):
            ref = cv2.distanceTransform(~masks[m], cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
            assert np.array_equal(got[m], ref), m
    finally:
        cv2.ipp.setUseIPP(ipp)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [synth.TUM_SHAPE, (61, 77), synth.KITTI_SHAPE])
def test_gpu_distance_transform(matcher, shape):
    masks = _dt_masks(shape, 6)
    assert np.array_equal(matcher.DistanceTransform(masks), oracle.distance_transform(masks))
