"""The hot path at BASELINE.json's sizes, checked through size-independent properties plus oracle spot checks
(the oracle alone would need minutes at these sizes): batches of 64 KITTI stereo frames, 20k map points per frame,
hundreds of keyframes x 2000 descriptors."""
import hashlib

import numpy as np
import pytest

import oracle
from object_slam_b200 import sharding, synth

from matcher_cases import MP_KEYS, bounds, map_case, oracle_map

pytestmark = pytest.mark.gpu


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_kitti_batch_of_64_stereo_frames(gpu):
    from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches
    F, shape = 64, synth.KITTI_SHAPE
    pairs = [synth.stereo_pair(shape, 1000 + i) for i in range(F)]
    eL = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=F)
    eR = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=F)
    runs = []
    for _ in range(2):
        resL = eL.extract_batch([p[0] for p in pairs])
        resR = eR.extract_batch([p[1] for p in pairs])
        st = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
        runs.append((resL, resR, st))
    # idempotence: the same batch twice gives the same bytes
    for a, b in zip(runs[0], runs[1]):
        assert _digest(*[x for r in a for x in r]) == _digest(*[x for r in b for x in r])
    resL, resR, st = runs[0]
    one = ORBextractor(2000, 1.2, 8, 20, 7)
    for i in (0, 17, 63):                                   # batch element == single call == oracle
        k, d = one(pairs[i][0])
        assert k.tobytes() == resL[i][0].tobytes() and np.array_equal(d, resL[i][1])
    ok, od = oracle.OracleExtractor(2000)(pairs[63][0])
    assert ok.tobytes() == resL[63][0].tobytes() and np.array_equal(od, resL[63][1])
    for (k, d), (ur, dp) in zip(resL, st):
        n = len(k)
        assert 1990 <= n <= 2024 and len(ur) == n
        assert np.all(np.diff(k["octave"]) >= 0)            # level-major order
        m = ur >= 0
        assert m.sum() > 300
        assert np.all(ur[m] <= k["x"][m] + 1e-3) and np.all(dp[m] > 0)       # disparity >= 0
        assert np.allclose(dp[m], synth.KITTI_BF / np.maximum(k["x"][m] - ur[m], 0.01), rtol=1e-5)


def test_projection_search_20k_points_128_frames(gpu):
    from object_slam_b200.matcher import ORBmatcher
    M = ORBmatcher(0.8, True)
    shape, B = synth.TUM_SHAPE, 128
    cases = [map_case(shape, 1000, 20000, 300 + s, 0.1) for s in range(4)]
    fs = M.frame_set(synth.scale_factors(), bounds(shape), synth.camera_for(shape), max_frames=B, max_keypoints=1000)
    fs.upload([cases[b % 4][0] for b in range(B)])
    arrs = [np.stack([cases[b % 4][1][k] for b in range(B)]) for k in MP_KEYS]
    kp_obs = np.zeros((B, fs.cap), np.int32)
    for b in range(B):
        kp_obs[b, :1000] = cases[b % 4][2]
    n, match = M.SearchByProjection(fs, *arrs, th=3.0, per_frame=True, kp_observations=kp_obs)
    n2, match2 = M.SearchByProjection(fs, *arrs, th=3.0, per_frame=True, kp_observations=kp_obs)
    assert np.array_equal(n, n2) and np.array_equal(match, match2)            # deterministic
    for b in range(B):                                                        # equal frames give equal results
        assert n[b] == n[b % 4] and np.array_equal(match[b], match[b % 4])
    for b in (0, 3):
        on, om = oracle_map(cases[b][0], shape, cases[b][1], 3.0, 0.8, cases[b][2])
        assert n[b] == on and np.array_equal(match[b, :1000], om)
    got = match >= 0
    assert not np.any(got & (kp_obs > 0))                                     # occupied keypoints stay untouched
    for b in range(4):
        mp = cases[b][1]
        idx = match[b][got[b]]
        assert np.all(mp["in_view"][idx] == 1)
        lock = mp["observations"][idx] > 0
        assert len(set(idx[lock])) == lock.sum()                              # a locking point owns at most one keypoint
    M.close()


def test_keyframe_matching_256_keyframes(gpu):
    from object_slam_b200.matcher import ORBmatcher
    M = ORBmatcher(0.6, True)
    K, n, W = 256, 2000, 8
    D = synth.keyframe_descriptors(K, n, 42)
    pairs = sharding.window_pairs(0, K, K, W)
    bi, bd, sd = M.knn2(D, pairs)
    bi2, bd2, sd2 = M.knn2(D, pairs)
    assert _digest(bi, bd, sd) == _digest(bi2, bd2, sd2)
    assert np.all(bd <= sd) and bd.min() >= 0 and sd.max() <= 256
    acc = bi >= 0
    assert np.all(bd[acc] <= 50) and np.all(bd[acc].astype(np.float32) < np.float32(0.6) * sd[acc].astype(np.float32))
    assert np.all(bi[~acc] == -1)
    # the accepted index really is at the reported distance (spot check with numpy popcounts)
    rng = np.random.default_rng(0)
    for p in rng.integers(0, len(pairs), 50):
        a, b = pairs[p]
        q = rng.integers(0, n, 40)
        full = np.unpackbits(D[a][q][:, None, :] ^ D[b][None, :, :], axis=-1).sum(-1)
        assert np.array_equal(full.min(1), bd[p][q])
        first = full.argmin(1)
        ok = bi[p][q] >= 0
        assert np.array_equal(first[ok], bi[p][q][ok])
    # planted near-duplicates between consecutive keyframes are found
    fwd = np.array([i for i, (a, b) in enumerate(pairs) if b == a - 1])
    assert (bi[fwd] >= 0).sum() > 0.2 * n * len(fwd)
    # sharding invariance: two ranks' halves concatenate to the single-process result
    parts = [sharding.window_pairs(*sharding.shard_range(K, r, 2), K, W) for r in range(2)]
    res = [M.knn2(D, p)[0] for p in parts]
    assert np.array_equal(np.concatenate(res), bi)
    o = oracle.hamming_knn2(D[pairs[100][0]], D[pairs[100][1]], 50, 0.6)
    assert np.array_equal(o[0], bi[100]) and np.array_equal(o[1], bd[100]) and np.array_equal(o[2], sd[100])
    M.close()
