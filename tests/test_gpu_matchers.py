"""Parity of the CUDA matchers (csrc/matcher.cu, through the C ABI) with the oracle: bit-exact match
indices, counts and updated state, on the same seeded inputs."""
import glob
import os

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth

from matcher_cases import MP_KEYS, LAST_KEYS, bounds, map_case, oracle_frame, oracle_init, oracle_last, oracle_map

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TUM, KITTI = synth.TUM_SHAPE, synth.KITTI_SHAPE


@pytest.fixture(scope="module")
def M(gpu):
    from object_slam_b200.matcher import ORBmatcher
    m = ORBmatcher(0.8, True)
    yield m
    m.close()


def frame_set(M, shape, frames, cap=None):
    cap = cap or max(len(f[0]) for f in frames)
    fs = M.frame_set(synth.scale_factors(), bounds(shape), synth.camera_for(shape), max_frames=len(frames), max_keypoints=max(cap, 1))
    return fs.upload(frames)


def test_descriptor_distance_and_three_maxima(M):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b[:100] = a[:100]
    got = M.DescriptorDistance(a, b)
    want = np.unpackbits(a ^ b, axis=1).sum(1)
    assert np.array_equal(got, want)
    h = rng.integers(0, 40, (500, 30)).astype(np.int32)
    h[:50] = 0
    h[50:100] = (rng.random((50, 30)) < 0.1) * rng.integers(0, 100, (50, 30))
    got = M.ComputeThreeMaxima(h)
    for i in range(len(h)):
        assert tuple(got[i]) == oracle.compute_three_maxima(h[i])


@pytest.mark.parametrize("shape,n", [(TUM, 1000), (KITTI, 2000), (TUM, 37), (TUM, 0)])
def test_grid_matches_assign_features_to_grid(M, shape, n):
    frames = [synth.synthetic_frame(shape, n, s) for s in (1, 2, 3)]
    if n:       # keypoints on and beyond the borders exercise the round()/drop rule of PosInGrid
        frames[0][0]["x"][:5] = [0.0, shape[1] - 0.4, shape[1] - 5.0, -3.0, shape[1] + 2.0]
        frames[0][0]["y"][5:9] = [0.0, shape[0] - 0.4, shape[0] - 5.1, -1.0]
    fs = frame_set(M, shape, frames, cap=max(n, 1))
    for b, f in enumerate(frames):
        start, idx = fs.grid(b)
        ostart, oidx = oracle_frame(f, shape).grid()
        assert np.array_equal(start, ostart) and np.array_equal(idx, oidx)


def run_map(M, shape, cases, th, nnratio, per_frame=True):
    M.mfNNratio = nnratio
    frames = [c[0] for c in cases]
    fs = frame_set(M, shape, frames)
    if per_frame:
        arrs = [np.stack([c[1][k] for c in cases]) for k in MP_KEYS]
    else:
        arrs = [cases[0][1][k] for k in MP_KEYS]
    kp_obs = None
    if any(c[2] is not None for c in cases):
        kp_obs = np.zeros((len(cases), fs.cap), np.int32)
        for b, c in enumerate(cases):
            if c[2] is not None:
                kp_obs[b, :len(c[2])] = c[2]
    n, match = M.SearchByProjection(fs, *arrs, th=th, per_frame=per_frame, kp_observations=kp_obs)
    return n, match, fs


@pytest.mark.parametrize("n_kp,n_mp,locked,th,anchored", [
    (1000, 20000, 0.0, 3.0, 0.5),      # cfg3: RGB-D tracking, th = 3 (Tracking.cc:1451)
    (1000, 3000, 0.25, 1.0, 0.7),      # th == 1: no factor (ORBmatcher.cc:49)
    (2000, 5000, 0.1, 5.0, 0.9),       # wide windows (th = 5 after relocalisation, Tracking.cc:1454): > 8 candidates per point
    (50, 400, 0.0, 3.0, 0.5),
])
def test_search_by_projection_matches_oracle(M, n_kp, n_mp, locked, th, anchored):
    shape = TUM
    cases = [map_case(shape, n_kp, n_mp, s, locked, anchored) for s in (0, 1, 2)]
    n, match, fs = run_map(M, shape, cases, th, 0.8)
    total = 0
    for b, (frame, mp, kp_obs) in enumerate(cases):
        on, omatch = oracle_map(frame, shape, mp, th, 0.8, kp_obs)
        assert n[b] == on
        assert np.array_equal(match[b, :n_kp], omatch)
        assert np.all(match[b, n_kp:] == -1)
        total += on
    assert total > 0


def test_search_by_projection_shared_points_and_edges(M):
    shape = TUM
    frame, mp, _ = map_case(shape, 800, 4000, 5)
    other = synth.synthetic_frame(shape, 640, 6)
    empty = synth.synthetic_frame(shape, 0, 7)
    cases = [(frame, mp, None), (other, mp, None), (empty, mp, None)]
    n, match, fs = run_map(M, shape, cases, 3.0, 0.8, per_frame=False)
    for b, (f, _, _) in enumerate(cases):
        on, omatch = oracle_map(f, shape, mp, 3.0, 0.8)
        assert n[b] == on and np.array_equal(match[b, :len(f[0])], omatch)
    assert n[2] == 0
    # nothing in view / every keypoint already taken / no points at all
    mp0 = dict(mp); mp0["in_view"] = np.zeros_like(mp["in_view"])
    n, match, _ = run_map(M, shape, [(frame, mp0, None)], 3.0, 0.8)
    assert n[0] == 0 and np.all(match == -1)
    n, match, _ = run_map(M, shape, [(frame, mp, np.ones(800, np.int32))], 3.0, 0.8)
    assert n[0] == 0 and np.all(match == -1)
    mpe = {k: v[:0] for k, v in mp.items()}
    n, match, _ = run_map(M, shape, [(frame, mpe, None)], 3.0, 0.8)
    assert n[0] == 0 and np.all(match == -1)
    # points without observations never lock: later points overwrite (n counts every assignment)
    mpz = dict(mp); mpz["observations"] = np.zeros_like(mp["observations"])
    n, match, _ = run_map(M, shape, [(frame, mpz, None)], 3.0, 0.8)
    on, omatch = oracle_map(frame, shape, mpz, 3.0, 0.8)
    assert n[0] == on and np.array_equal(match[0, :800], omatch) and on > (omatch >= 0).sum()
    assert M.last_rounds()[0] == (4000 + 1023) // 1024        # no locking point: one pass per block of 1024 points


def test_search_by_projection_adversarial_chain(M):
    """Every point wants the same few keypoints: long dependency chains for the lock resolution."""
    shape = TUM
    frame = synth.synthetic_frame(shape, 64, 9)
    keys, desc, ur = frame
    keys["x"] = 300 + (np.arange(64) % 8) * 3.0
    keys["y"] = 200 + (np.arange(64) // 8) * 3.0
    keys["octave"] = 0
    ur[:] = -1
    rng = np.random.default_rng(10)
    n_mp = 3000
    mp = dict(in_view=np.ones(n_mp, np.uint8), proj_x=np.full(n_mp, 310, np.float32), proj_y=np.full(n_mp, 210, np.float32),
              proj_xr=np.zeros(n_mp, np.float32), scale_level=np.zeros(n_mp, np.int32), view_cos=np.full(n_mp, 0.9, np.float32),
              descriptors=synth.flip_bits(desc[rng.integers(0, 64, n_mp)], rng.integers(0, 90, n_mp), rng),
              observations=rng.integers(0, 2, n_mp).astype(np.int32))
    n, match, _ = run_map(M, shape, [(frame, mp, None)], 3.0, 0.95)
    on, omatch = oracle_map(frame, shape, mp, 3.0, 0.95)
    assert n[0] == on and np.array_equal(match[0, :64], omatch)
    assert M.last_rounds()[0] >= 3


@pytest.mark.parametrize("mono,forward,th", [(False, 0.0, 7.0), (False, 0.6, 7.0), (False, -0.6, 14.0), (True, 0.6, 15.0)])
def test_search_by_projection_last_matches_oracle(M, mono, forward, th):
    shape = TUM
    pairs = [synth.motion_pair(shape, 1000, s, forward=forward) for s in (20, 21, 22)]
    fs = frame_set(M, shape, [p[1] for p in pairs])
    arrs = [np.stack([p[0][k] for p in pairs]) for k in LAST_KEYS]
    tl = np.stack([p[0]["tcw_last"] for p in pairs]); tc = np.stack([p[0]["tcw_current"] for p in pairs])
    for check_ori in (True, False):
        M.mbCheckOrientation = check_ori
        n, match = M.SearchByProjectionLast(fs, *arrs, tl, tc, th, mono, per_frame=True)
        for b, (last, cur) in enumerate(pairs):
            on, omatch = oracle_last(cur, shape, last, th, mono, check_ori)
            assert n[b] == on and np.array_equal(match[b, :1000], omatch)
            assert on > 200
    M.mbCheckOrientation = True


@pytest.mark.parametrize("shape,n,window,ratio", [(KITTI, 2000, 100, 0.9), (TUM, 1000, 100, 0.9), (TUM, 1000, 30, 0.6), (TUM, 5, 100, 0.9)])
def test_search_for_initialization_matches_oracle(M, shape, n, window, ratio):
    M.mfNNratio = ratio
    cases = [synth.init_pair(shape, n, s) for s in (40, 41)]
    f1 = frame_set(M, shape, [c[0] for c in cases])
    f2 = frame_set(M, shape, [c[1] for c in cases])
    prev = np.zeros((2, f1.cap, 2), np.float32)
    for b, c in enumerate(cases):
        prev[b, :n] = c[2]
    for check_ori in (True, False):
        M.mbCheckOrientation = check_ori
        pv = prev.copy()
        nm, m12 = M.SearchForInitialization(f1, f2, pv, window)
        for b, (a, bb, p) in enumerate(cases):
            on, om12, opm = oracle_init(a, bb, shape, p, window, ratio, check_ori)
            assert nm[b] == on and np.array_equal(m12[b, :n], om12) and np.array_equal(pv[b, :n], opm)
    # second round, as Tracking calls it again with the updated vbPrevMatched
    nm2, m12b = M.SearchForInitialization(f1, f2, pv, window)
    on, om12, opm = oracle_init(cases[0][0], cases[0][1], shape, oracle_init(cases[0][0], cases[0][1], shape, cases[0][2], window, ratio, False)[2], window, ratio, False)
    assert nm2[0] == on and np.array_equal(m12b[0, :n], om12)
    M.mbCheckOrientation = True
    M.mfNNratio = 0.8


def test_knn2_matches_oracle(M):
    M.mfNNratio = 0.6
    D = synth.keyframe_descriptors(5, 2000, 3)
    pairs = np.array([(i, j) for i in range(5) for j in range(5) if i != j], np.int32)
    bi, bd, sd = M.knn2(D, pairs)
    for p, (a, b) in enumerate(pairs):
        obi, obd, osd = oracle.hamming_knn2(D[a], D[b], 50, 0.6)
        assert np.array_equal(bi[p], obi) and np.array_equal(bd[p], obd) and np.array_equal(sd[p], osd)
    assert (bi >= 0).sum() > 1000
    # ragged size (not a multiple of the tile), self match, duplicate database rows: lowest index wins ties
    D2 = synth.keyframe_descriptors(2, 333, 4)
    D2[1, 100] = D2[1, 7]
    D2[0, 5] = D2[1, 7]
    bi, bd, sd = M.knn2(D2, np.array([(0, 1), (1, 1)], np.int32))
    for p, (a, b) in enumerate([(0, 1), (1, 1)]):
        obi, obd, osd = oracle.hamming_knn2(D2[a], D2[b], 50, 0.6)
        assert np.array_equal(bi[p], obi) and np.array_equal(bd[p], obd) and np.array_equal(sd[p], osd)
    assert bd[0, 5] == 0 and sd[0, 5] == 0 and bi[0, 5] == -1
    M.mfNNratio = 0.8


@pytest.mark.parametrize("engine", ["tensor", "tensor_single_cta", "popc"])
def test_knn2_engines_match_oracle(M, engine):
    """The tcgen05 int8 contraction (hamming = (256 - a.b) / 2) and the POPC kernel against the oracle: sizes that are not
    multiples of the 128 x 256 tile, a size below one tile, duplicates (lowest index wins, second best counts duplicates),
    complementary descriptors (distance 256), an all-equal keyframe."""
    from object_slam_b200._capi import lib, check
    M.set_knn2_engine(M.KNN2_POPC if engine == "popc" else M.KNN2_TENSOR)
    # "tensor" = CTA pairs (tcgen05.mma.cta_group::2, the default), "tensor_single_cta" = one CTA per SM
    check(lib().obs_set_option(b"knn2_cta_pair", 0 if engine == "tensor_single_cta" else 1))
    M.mfNNratio = 0.6
    try:
        for n, seed in ((2000, 11), (333, 12), (129, 13), (257, 14), (70, 15), (1, 16), (1024, 17), (1900, 20)):
            D = synth.keyframe_descriptors(3, n, seed)
            if n >= 70:
                D[1, 40] = D[1, 7]                  # duplicate database rows
                D[0, 5] = D[1, 7]                   # ... matched exactly: best = second = 0
                D[2, 3] = ~D[1, 9]                  # complement: distance 256
                D[2, n - 1] = D[1, n - 1]           # last row / last column of the ragged tiles
            pairs = np.array([(0, 1), (1, 0), (2, 1), (1, 1), (2, 2)], np.int32)
            bi, bd, sd = M.knn2(D, pairs)
            for p, (a, b) in enumerate(pairs):
                obi, obd, osd = oracle.hamming_knn2(D[a], D[b], 50, 0.6)
                assert np.array_equal(bd[p], obd), (engine, n, p)
                assert np.array_equal(sd[p], osd), (engine, n, p)
                assert np.array_equal(bi[p], obi), (engine, n, p)
            if n >= 70:
                assert bd[0, 5] == 0 and sd[0, 5] == 0 and bi[0, 5] == -1 and bd[2, n - 1] == 0
        # every descriptor equal: all distances 0, the first database row wins everywhere
        D = np.repeat(synth.keyframe_descriptors(1, 1, 18), 300, axis=1).repeat(2, axis=0)
        bi, bd, sd = M.knn2(D, np.array([(0, 1)], np.int32))
        assert (bd == 0).all() and (sd == 0).all() and (bi == -1).all()
        # many pairs: more work items than SMs, every accumulator / stage parity exercised
        D = synth.keyframe_descriptors(12, 600, 19)
        pairs = np.array([(i, j) for i in range(12) for j in range(12) if abs(i - j) in (1, 2, 3)], np.int32)
        bi, bd, sd = M.knn2(D, pairs)
        for p, (a, b) in enumerate(pairs):
            obi, obd, osd = oracle.hamming_knn2(D[a], D[b], 50, 0.6)
            assert np.array_equal(bi[p], obi) and np.array_equal(bd[p], obd) and np.array_equal(sd[p], osd), (engine, p)
    finally:
        M.set_knn2_engine(M.KNN2_AUTO)
        check(lib().obs_set_option(b"knn2_cta_pair", 1))
        M.mfNNratio = 0.8


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "match_*.npz"))), ids=os.path.basename)
def test_golden_fixtures(M, path):
    g = np.load(path)
    kind, seed = str(g["kind"]), int(g["seed"])
    shape = TUM
    if kind == "map":
        case = map_case(shape, 1000, 20000 if seed == 0 else 3000, seed, 0.0 if seed == 0 else 0.25)
        n, match, _ = run_map(M, shape, [case], 3.0, 0.8)
        assert n[0] == g["n_matches"] and np.array_equal(match[0, :1000], g["kp_match"])
    elif kind in ("last", "last_forward"):
        last, cur = synth.motion_pair(shape, 1000, seed, forward=0.6 if kind == "last_forward" else 0.0)
        fs = frame_set(M, shape, [cur])
        n, match = M.SearchByProjectionLast(fs, *[last[k] for k in LAST_KEYS], last["tcw_last"], last["tcw_current"], 7.0, False)
        m = match[0, :1000]
        assert n[0] == g["n_matches"] and np.array_equal(np.where(m == -2, -1, m), g["kp_match"])     # -2: NULL again after the rotation check
    elif kind in ("keyframe", "sim3"):
        last, cur = synth.motion_pair(shape, 1000, seed)
        pts = synth.keyframe_points(last, seed + 100)
        fs = frame_set(M, shape, [cur])
        if kind == "keyframe":
            n, match = M.SearchByProjectionKeyFrame(fs, pts, last["tcw_current"], 10.0, 100)
        else:
            # the reference decomposes Scw itself (ORBmatcher.cc:299-303); the fixture carries its [Rcw | tcw]
            n, match = M.SearchByProjectionSim3(fs, pts, g["rt_decomposed"], 10)
        m = match[0, :1000]
        assert n[0] == g["n_matches"] and np.array_equal(np.where(m == -2, -1, m), g["kp_match"])
    elif kind == "init":
        M.mfNNratio = 0.9
        f1, f2, prev = synth.init_pair(KITTI, 2000, seed)
        s1, s2 = frame_set(M, KITTI, [f1]), frame_set(M, KITTI, [f2])
        pv = np.zeros((1, s1.cap, 2), np.float32); pv[0, :2000] = prev
        n, m12 = M.SearchForInitialization(s1, s2, pv, 100)
        M.mfNNratio = 0.8
        assert n[0] == g["n_matches"] and np.array_equal(m12[0, :2000], g["matches12"]) and np.array_equal(pv[0, :2000], g["prev_matched"])
    elif kind == "knn2":
        M.mfNNratio = 0.6
        D = synth.keyframe_descriptors(3, 2000, seed)
        bi, bd, sd = M.knn2(D, np.array([(1, 0)], np.int32))
        M.mfNNratio = 0.8
        assert np.array_equal(bi[0], g["best_idx"]) and np.array_equal(bd[0], g["best_dist"]) and np.array_equal(sd[0], g["second_dist"])


def test_device_pointers_and_extractor_frames(M):
    """Device-resident inputs/outputs (torch tensors) and frames taken straight from an extraction."""
    import torch
    from object_slam_b200.extractor import ORBextractor
    shape = TUM
    imgs = [synth.blocky_image(shape, s) for s in (50, 51)]
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_size=(640, 480), max_batch=2)
    res = ex.extract_batch(imgs)
    fs = M.frame_set(ex.GetScaleFactors(), bounds(shape), synth.camera_for(shape), max_frames=2, max_keypoints=ex.capacity)
    fs.from_extractor(ex)
    mps = [synth.map_points_for_frame(k, d, shape, 6000, 60 + b) for b, (k, d) in enumerate(res)]
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.stack([m[k] for m in mps])).to(dev) for k in MP_KEYS]
    kp_match = torch.full((2, fs.cap), -7, dtype=torch.int32, device=dev)
    n_matches = torch.zeros(2, dtype=torch.int32, device=dev)
    M.mfNNratio = 0.8
    torch.cuda.synchronize()
    M.SearchByProjection(fs, *[x.data_ptr() for x in t], th=3.0, n_points=6000, per_frame=True,
                         kp_match=kp_match.data_ptr(), n_matches=n_matches.data_ptr())
    M.sync()
    for b, (k, d) in enumerate(res):
        start, idx = fs.grid(b)
        ostart, oidx = oracle_frame((k, d, None), shape).grid()
        assert np.array_equal(start, ostart) and np.array_equal(idx, oidx)
        on, omatch = oracle_map((k, d, None), shape, mps[b], 3.0, 0.8)
        assert int(n_matches[b]) == on and np.array_equal(kp_match[b, :len(k)].cpu().numpy(), omatch)
        assert on > 100


def test_errors(M):
    from object_slam_b200._capi import ObsError
    with pytest.raises(ObsError):
        M.frame_set(synth.scale_factors(), (0, 0, 0, 480), max_frames=1, max_keypoints=10)
    fs = M.frame_set(synth.scale_factors(), bounds(TUM), max_frames=1, max_keypoints=32)
    with pytest.raises(ObsError):
        fs.upload([synth.synthetic_frame(TUM, 100, 0)])          # more keypoints than the set holds
    with pytest.raises(ObsError):
        M.SearchByProjection(fs, *[np.zeros(1, np.uint8)] * 8)   # empty frame set


def test_capacity_errors_are_explicit(M):
    """Frame sets beyond what the searches can hold in shared memory fail with OBS_ERR_CAPACITY and a message, not with a
    launch error."""
    from object_slam_b200._capi import ObsError
    fs = M.frame_set(synth.scale_factors(), bounds(TUM), max_frames=1, max_keypoints=20000)
    f = synth.synthetic_frame(TUM, 300, 3)
    fs.upload([f])
    mp = synth.map_points_for_frame(f[0], f[1], TUM, 100, 1)
    with pytest.raises(ObsError, match="max_keypoints"):
        M.SearchByProjection(fs, *[mp[k] for k in MP_KEYS], th=3.0)
    with pytest.raises(ObsError, match="max_keypoints"):
        M.SearchForInitialization(fs, fs, np.zeros((1, fs.cap, 2), np.float32), 100)


def test_frame_set_searched_from_another_matcher(M):
    """A frame uploaded through one matcher (the Tracking thread's) is searched from another one (LocalMapping's): the second
    matcher's stream is ordered behind the build."""
    from object_slam_b200.matcher import ORBmatcher
    shape = TUM
    frame, mp, _ = map_case(shape, 800, 4000, 5)
    other = ORBmatcher(0.8, True)
    for _ in range(3):
        fs = frame_set(M, shape, [frame])
        n, match = other.SearchByProjection(fs, *[mp[k] for k in MP_KEYS], th=3.0, per_frame=False)
        on, omatch = oracle_map(frame, shape, mp, 3.0, 0.8)
        assert n[0] == on and np.array_equal(match[0, :800], omatch)


def test_two_gpu_allgather_and_sharded_matching(gpu):
    """The NCCL exchange step (csrc/comm.cu) and sharded keyframe matching on 2 GPUs; skipped on a 1-GPU box."""
    import subprocess
    import sys
    from object_slam_b200 import _capi
    if _capi.lib().obs_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "check_allgather.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


@pytest.mark.parametrize("th,orb_dist,forward", [(10.0, 100, 0.0), (3.0, 64, 0.0), (10.0, 100, 3.0)])
def test_search_by_projection_keyframe_matches_oracle(M, th, orb_dist, forward):
    """Relocalisation search (ORBmatcher.cc:1472-1599): projection, distance gate, PredictScale (device logf), rotation check."""
    shape = TUM
    pairs = [synth.motion_pair(shape, 1000, s, forward=forward) for s in (70, 71)]
    pts = [synth.keyframe_points(p[0], 80 + i) for i, p in enumerate(pairs)]
    fs = frame_set(M, shape, [p[1] for p in pairs])
    stacked = {k: np.stack([q[k] for q in pts]) for k in pts[0]}
    tcw = np.stack([p[0]["tcw_current"] for p in pairs])
    rng = np.random.default_rng(5)
    taken = (rng.random((2, fs.cap)) < 0.1).astype(np.int32)
    for check_ori in (True, False):
        M.mbCheckOrientation = check_ori
        n, match = M.SearchByProjectionKeyFrame(fs, stacked, tcw, th, orb_dist, per_frame=True, kp_taken=taken)
        for b, (last, cur) in enumerate(pairs):
            on, om = oracle.search_by_projection_keyframe(oracle_frame(cur, shape), synth.scale_factors(), synth.camera_for(shape),
                                                          last["tcw_current"], pts[b], th, orb_dist, check_ori, taken[b, :1000])
            assert n[b] == on and np.array_equal(match[b, :1000], om)
            assert on > 100 or forward > 0
    M.mbCheckOrientation = True


@pytest.mark.parametrize("th", [10, 4])
def test_search_by_projection_sim3_matches_oracle(M, th):
    """Loop-closing search (ORBmatcher.cc:290-403) after the decomposition of Scw."""
    shape = TUM
    pairs = [synth.motion_pair(shape, 1000, s) for s in (90, 91, 92)]
    pts = [synth.keyframe_points(p[0], 95 + i) for i, p in enumerate(pairs)]
    fs = frame_set(M, shape, [p[1] for p in pairs])
    stacked = {k: np.stack([q[k] for q in pts]) for k in pts[0]}
    tcw = np.stack([p[0]["tcw_current"] for p in pairs])
    rng = np.random.default_rng(6)
    taken = (rng.random((3, fs.cap)) < 0.2).astype(np.int32)
    n, match = M.SearchByProjectionSim3(fs, stacked, tcw, th, per_frame=True, kp_taken=taken)
    for b, (last, cur) in enumerate(pairs):
        on, om = oracle.search_by_projection_sim3(oracle_frame(cur, shape), synth.scale_factors(), synth.camera_for(shape),
                                                  last["tcw_current"], pts[b], th, taken[b, :1000])
        assert n[b] == on and np.array_equal(match[b, :1000], om)
        assert on > 100


@pytest.mark.parametrize("sim3,th", [(False, 3.0), (True, 4.0), (False, 2.5)])
def test_fuse_search_matches_oracle(M, sim3, th):
    """Search half of ORBmatcher::Fuse (ORBmatcher.cc:825-966 keyframe variant with the chi-square gates, :974-1100 Sim3 variant)."""
    shape = TUM
    pairs = [synth.motion_pair(shape, 1000, s) for s in (70, 71, 72)]
    pts = [synth.keyframe_points(p[0], 75 + i) for i, p in enumerate(pairs)]
    fs = frame_set(M, shape, [p[1] for p in pairs])
    stacked = {k: np.stack([q[k] for q in pts]) for k in pts[0]}
    tcw = np.stack([p[0]["tcw_current"] for p in pairs])
    ow = np.stack([oracle.minus_rt_t(t) for t in tcw])
    bi, bd = M.FuseSearch(fs, stacked, tcw, th, camera_centre=None if sim3 else ow, sim3=sim3, per_frame=True)
    total = 0
    for b, (last, cur) in enumerate(pairs):
        oi, od = oracle.fuse_search(oracle_frame(cur, shape), synth.scale_factors(), synth.camera_for(shape), last["tcw_current"], pts[b], th,
                                    camera_centre=ow[b], sim3=sim3)
        assert np.array_equal(bi[b], oi) and np.array_equal(bd[b], od)
        total += int(((od <= 50) & (oi >= 0)).sum())
    assert total > 100


@pytest.mark.parametrize("th", [7.5, 3.0])
def test_search_by_sim3_matches_oracle(M, th):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1102-1326): both projection directions and the agreement check."""
    shape = TUM
    cases = [synth.sim3_pair(shape, 1000, s) for s in (40, 41, 42)]
    s1 = frame_set(M, shape, [c[0] for c in cases]); s2 = frame_set(M, shape, [c[1] for c in cases])
    st = lambda i: {k: np.stack([c[i][k] for c in cases]) for k in cases[0][i]}
    T = {k: np.stack([c[4][k] for c in cases]) for k in cases[0][4]}
    nf, m12 = M.SearchBySim3(s1, s2, st(2), st(3), T["t1w"], T["t2w"], T["t21"], T["t12"], th, per_frame=True)
    for b, (k1, k2, p1, p2, P) in enumerate(cases):
        on, om = oracle.search_by_sim3(oracle_frame(k1, shape), oracle_frame(k2, shape), synth.scale_factors(), synth.camera_for(shape),
                                       P["t1w"], P["t2w"], P["t21"], P["t12"], p1, p2, th)
        assert nf[b] == on and np.array_equal(m12[b], om)
        assert on > 100
