"""The restated stereo matcher and matchers (oracle/match_oracle.cpp) against the reference's OWN object code: oracle/_ref/
libref_matcher.so is the unmodified src/ORBmatcher.cc, src/Frame.cc, src/MapPoint.cc and src/KeyFrame.cc of /root/reference compiled
in place against oracle/refstubs (oracle/Makefile).  Every comparison is bit for bit on seeded inputs; this is what pins rows
a11-a16 and f1-f2 of SURVEY section 8 (the golden fixtures under tests/golden are generated from the same library,
tools/make_golden_match.py).  Skipped where the library has not been built (it needs /root/reference; the built file travels)."""
import numpy as np
import pytest

import oracle
from oracle import refm
from object_slam_b200 import synth
from matcher_cases import LAST_KEYS, MP_KEYS, bounds, map_case, oracle_frame, oracle_init, oracle_last, oracle_map

pytestmark = pytest.mark.skipif(not refm.available(), reason="oracle/_ref/libref_matcher.so not built (needs /root/reference)")
TUM, KITTI = synth.TUM_SHAPE, synth.KITTI_SHAPE
SF = synth.scale_factors()


def ref_frame(frame, shape, tcw=None, cam=None):
    return refm.frame(frame[0], frame[1], frame[2], bounds(shape), synth.camera_for(shape) if cam is None else cam, SF, tcw)


def test_descriptor_distance_and_three_maxima():
    rng = np.random.default_rng(0)
    for _ in range(500):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert refm.descriptor_distance(a, b) == oracle.descriptor_distance(a, b)
    for _ in range(500):
        s = rng.integers(0, int(rng.integers(1, 25)), 30)
        assert refm.compute_three_maxima(s) == oracle.compute_three_maxima(s)
    assert refm.compute_three_maxima(np.zeros(30, np.int32)) == oracle.compute_three_maxima(np.zeros(30, np.int32))


@pytest.mark.parametrize("shape,n,seed", [(TUM, 1000, 3), (KITTI, 2000, 4), (TUM, 0, 5), (TUM, 1, 6)])
def test_grid_and_features_in_area(shape, n, seed):
    fr = synth.synthetic_frame(shape, n, seed)
    if n:      # keypoints on and beyond the image border: PosInGrid drops column 64 / row 48
        fr[0]["x"][0] = shape[1] - 0.01
        fr[0]["y"][0] = shape[0] - 0.01
    F, R = oracle_frame(fr, shape), ref_frame(fr, shape)
    s1, i1 = F.grid(); s2, i2 = refm.frame_grid(R)
    assert np.array_equal(s1, s2) and np.array_equal(i1, i2)
    rng = np.random.default_rng(seed)
    for _ in range(300):
        x, y, r = rng.uniform(-20, shape[1] + 20), rng.uniform(-20, shape[0] + 20), rng.uniform(1, 80)
        lo, hi = int(rng.integers(-1, 8)), int(rng.integers(-1, 8))
        assert np.array_equal(F.features_in_area(x, y, r, lo, hi), refm.features_in_area(R, x, y, r, lo, hi))
        assert np.array_equal(F.features_in_area(x, y, r, -1, -1), refm.features_in_area(R, x, y, r, keyframe=True))


@pytest.mark.parametrize("shape,nf,mbf,maxD,seeds", [(KITTI, 2000, synth.KITTI_BF, synth.KITTI_FX, (0, 1, 2, 3)), (TUM, 1000, 40.0, 525.0, (4, 5)),
                                                     (KITTI, 2000, synth.KITTI_BF, 30.0, (6,)), (KITTI, 500, synth.KITTI_BF, 1e-3, (7,))])
def test_compute_stereo_matches(shape, nf, mbf, maxD, seeds):
    total = 0
    for seed in seeds:
        L, R = synth.stereo_pair(shape, seed)
        oL, oR = oracle.OracleExtractor(nf), oracle.OracleExtractor(nf)
        kL, dL = oL(L); kR, dR = oR(R)
        t = oL.tables()
        pl = [oL.level(l) for l in range(8)]; pr = [oR.level(l) for l in range(8)]
        ur, dp, used = refm.stereo_match(kL, dL, kR, dR, pl, pr, t["scale"], mbf, maxD, bounds(shape))
        our, odp, _ = oracle.stereo_match(kL, dL, kR, dR, pl, pr, t["scale"], t["inv_scale"], mbf, 0.0, used)
        assert np.array_equal(ur, our) and np.array_equal(dp, odp)
        total += int((ur >= 0).sum())
    assert total > 100 or maxD < 1
    # unrelated right image: (almost) nothing matches
    Rn = synth.noise_image(shape, 1)
    kR, dR = oR(Rn); pr = [oR.level(l) for l in range(8)]
    ur, dp, used = refm.stereo_match(kL, dL, kR, dR, pl, pr, t["scale"], mbf, maxD, bounds(shape))
    our, odp, _ = oracle.stereo_match(kL, dL, kR, dR, pl, pr, t["scale"], t["inv_scale"], mbf, 0.0, used)
    assert np.array_equal(ur, our) and np.array_equal(dp, odp)


@pytest.mark.parametrize("n_kp,n_mp,locked,th,anchored,seeds", [(1000, 20000, 0.0, 3.0, 0.5, (0,)), (1000, 3000, 0.25, 1.0, 0.7, (1, 2)),
                                                                (2000, 5000, 0.1, 5.0, 0.9, (3,)), (50, 400, 0.0, 3.0, 0.5, (4, 5, 6)),
                                                                (800, 4000, 0.5, 3.0, 0.95, (7,))])
def test_search_by_projection_map(n_kp, n_mp, locked, th, anchored, seeds):
    total = 0
    for seed in seeds:
        frame, mp, kp_obs = map_case(TUM, n_kp, n_mp, seed, locked, anchored)
        for ratio in (0.8, 0.6):
            n, match = oracle_map(frame, TUM, mp, th, ratio, kp_obs)
            rn, rmatch = refm.search_by_projection_map(ref_frame(frame, TUM), *[mp[k] for k in MP_KEYS], th, ratio, kp_obs)
            assert n == rn and np.array_equal(match, rmatch)
            total += n
    assert total > 0
    # points without observations never lock their keypoint: later points overwrite
    mpz = dict(mp); mpz["observations"] = np.zeros_like(mp["observations"])
    n, match = oracle_map(frame, TUM, mpz, th, 0.8)
    rn, rmatch = refm.search_by_projection_map(ref_frame(frame, TUM), *[mpz[k] for k in MP_KEYS], th, 0.8)
    assert n == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("mono,forward,th,seed", [(False, 0.0, 7.0, 0), (True, 0.0, 15.0, 1), (False, 0.6, 7.0, 2), (False, -0.6, 7.0, 3), (True, 0.4, 7.0, 4)])
def test_search_by_projection_last(mono, forward, th, seed):
    last, cur = synth.motion_pair(TUM, 1000, seed, forward=forward)
    for check_ori in (True, False):
        rng = np.random.default_rng(seed)
        kp_obs = (rng.random(1000) < 0.1).astype(np.int32) if check_ori else None
        n, match = oracle_last(cur, TUM, last, th, mono, check_ori, kp_obs)
        R = ref_frame(cur, TUM, last["tcw_current"])
        rn, rmatch = refm.search_by_projection_last(R, last["tcw_last"], *[last[k] for k in LAST_KEYS], th, mono, check_ori, kp_obs)
        assert n == rn and n > 100
        assert np.array_equal(np.where(match == -2, -1, match), rmatch)       # -2 marks matches the rotation check removed


@pytest.mark.parametrize("shape,n,window,ratio,seed", [(TUM, 1000, 100, 0.9, 0), (TUM, 2000, 100, 0.9, 1), (KITTI, 2000, 50, 0.8, 2), (TUM, 300, 10, 0.9, 3)])
def test_search_for_initialization(shape, n, window, ratio, seed):
    f1, f2, prev = synth.init_pair(shape, n, seed)
    for check_ori in (True, False):
        on, om12, opm = oracle_init(f1, f2, shape, prev, window, ratio, check_ori)
        rn, rm12, rpm = refm.search_for_initialization(ref_frame(f1, shape), ref_frame(f2, shape), prev, window, ratio, check_ori)
        assert on == rn and np.array_equal(om12, rm12) and np.array_equal(opm, rpm)
    assert on > 10 or window < 20


@pytest.mark.parametrize("th,orb_dist,forward,seed", [(10.0, 100, 0.0, 5), (3.0, 64, 0.0, 6), (10.0, 100, 3.0, 7), (15.0, 100, -1.0, 8)])
def test_search_by_projection_keyframe(th, orb_dist, forward, seed):
    last, cur = synth.motion_pair(TUM, 1000, seed, forward=forward)
    pts = refm.canonical_points(synth.keyframe_points(last, seed + 100))
    cam = synth.camera_for(TUM)
    rng = np.random.default_rng(seed)
    for check_ori, taken in ((True, None), (False, (rng.random(1000) < 0.2).astype(np.int32))):
        n, m = oracle.search_by_projection_keyframe(oracle_frame(cur, TUM), SF, cam, last["tcw_current"], pts, th, orb_dist, check_ori, taken)
        rn, rm = refm.search_by_projection_keyframe(ref_frame(cur, TUM, last["tcw_current"]), pts, th, orb_dist, check_ori, taken)
        assert n == rn and np.array_equal(np.where(m == -2, -1, m), rm)
    assert n > 20


def test_predict_scale():
    rng = np.random.default_rng(1)
    for _ in range(3000):
        raw, d = np.float32(10 ** rng.uniform(-1, 2)), np.float32(10 ** rng.uniform(-1, 2))
        assert refm.predict_scale(raw, d, SF) == oracle.lib().orc_predict_scale(oracle.C.c_float(raw), oracle.C.c_float(d), oracle.C.c_float(oracle.logf(SF[1])), 8)


@pytest.mark.parametrize("th,seed,scale", [(10, 5, 1.0), (3, 6, 1.0), (10, 7, 1.37)])
def test_search_by_projection_sim3(th, seed, scale):
    """SearchByProjection(KeyFrame*, Scw, ...): the reference decomposes Scw itself (:299-303); the restatement starts behind that."""
    last, cur = synth.motion_pair(TUM, 1000, seed)
    pts = refm.canonical_points(synth.keyframe_points(last, seed + 100))
    cam = synth.camera_for(TUM)
    scw = np.array(last["tcw_current"], np.float32).reshape(3, 4).copy()
    scw[:, :3] *= np.float32(scale); scw[:, 3] *= np.float32(scale)
    rt, ow = refm.decompose_scw(scw)
    rng = np.random.default_rng(seed)
    for taken in (None, (rng.random(1000) < 0.3).astype(np.int32)):
        n, m = oracle.search_by_projection_sim3(oracle_frame(cur, TUM), SF, cam, rt, pts, th, taken)
        rn, rm = refm.search_by_projection_sim3(ref_frame(cur, TUM), scw, pts, th, taken)
        assert n == rn and np.array_equal(m, rm)
    assert n > 30


@pytest.mark.parametrize("n,seed,nodes", [(500, 1, 40), (2000, 2, 100), (300, 3, 1), (60, 4, 400)])
def test_search_by_bow(n, seed, nodes):
    a, b, _ = synth.bow_pair(TUM, n, seed, n_nodes=nodes)
    for ratio, ori in ((0.7, True), (0.9, False)):
        on, om12, om21 = oracle.search_by_bow(a, dict(b, valid=None), 50, False, ratio, ori)           # SearchByBoW(KeyFrame*, Frame&): TH_LOW
        rn, rm12, rm21 = refm.search_by_bow(a, dict(b, valid=None), bounds(TUM), False, ratio, ori)
        assert on == rn and np.array_equal(om21, rm21)
        on, om12, om21 = oracle.search_by_bow(a, b, 50, True, ratio, ori)                               # SearchByBoW(KeyFrame*, KeyFrame*)
        rn, rm12, rm21 = refm.search_by_bow(a, b, bounds(TUM), True, ratio, ori)
        assert on == rn and np.array_equal(om12, rm12)
    assert on > 5 or n < 100


@pytest.mark.parametrize("n,seed,only_stereo", [(500, 11, False), (2000, 12, False), (800, 13, True)])
def test_search_for_triangulation(n, seed, only_stereo):
    a, b, extra = synth.bow_pair(TUM, n, seed, n_nodes=60)
    cam = synth.camera_for(TUM)
    rng = np.random.default_rng(seed)
    t1 = np.hstack([synth._rot(*rng.normal(0, 0.02, 3)), rng.normal(0, 0.3, (3, 1))]).astype(np.float32)
    t2 = np.hstack([synth._rot(*rng.normal(0, 0.02, 3)), rng.normal(0, 0.3, (3, 1))]).astype(np.float32)
    f12 = np.asarray(extra["f12"], np.float32) if isinstance(extra, dict) and "f12" in extra else (rng.normal(0, 1, (3, 3)) * 1e-3).astype(np.float32)
    s2 = (SF * SF).astype(np.float32)
    for ori in (True, False):
        rn, rm12, ep = refm.search_for_triangulation(a, b, f12, bounds(TUM), cam, SF, t1, t2, only_stereo, ori)
        on, om12 = oracle.search_for_triangulation(a, b, f12, ep, s2, SF, only_stereo, ori)
        assert on == rn and np.array_equal(om12, rm12)


@pytest.mark.parametrize("sim3,th,seed", [(False, 3.0, 20), (False, 2.5, 21), (True, 4.0, 22), (True, 3.0, 23)])
def test_fuse(sim3, th, seed):
    last, cur = synth.motion_pair(TUM, 1000, seed)
    pts = refm.canonical_points(synth.keyframe_points(last, seed + 100))
    cam = synth.camera_for(TUM)
    tcw = np.array(last["tcw_current"], np.float32).reshape(3, 4)
    if sim3:
        scw = tcw.copy()
        rt, ow = refm.decompose_scw(scw)
        bi, bd = oracle.fuse_search(oracle_frame(cur, TUM), SF, cam, rt, pts, th, ow, True)
        rbi, total = refm.fuse(ref_frame(cur, TUM, tcw), pts, th, scw)
    else:
        bi, bd = oracle.fuse_search(oracle_frame(cur, TUM), SF, cam, tcw, pts, th, None, False)
        rbi, total = refm.fuse(ref_frame(cur, TUM, tcw), pts, th)
    want = np.where(bd <= 50, bi, -1)
    assert np.array_equal(want, rbi)
    assert total == (rbi >= 0).sum() and total > 50


@pytest.mark.parametrize("th,seed", [(7.5, 30), (4.0, 31), (10.0, 32)])
def test_search_by_sim3(th, seed):
    kf1, kf2, pts1, pts2, poses = synth.sim3_pair(TUM, 800, seed)
    pts1, pts2 = refm.canonical_points(pts1), refm.canonical_points(pts2)
    cam = synth.camera_for(TUM)
    rng = np.random.default_rng(seed)
    # the reference receives (s12, R12, t12) and builds [sR21 | t21], [sR12 | t12] itself (:1119-1122)
    s12 = np.float32(1.0 + rng.normal(0, 0.01))
    r12 = (np.asarray(poses["t12"], np.float32).reshape(3, 4)[:, :3] / np.float32(np.linalg.norm(np.asarray(poses["t12"]).reshape(3, 4)[0, :3]))).astype(np.float32)
    t12 = np.asarray(poses["t12"], np.float32).reshape(3, 4)[:, 3].copy()
    T21, T12 = refm.sim3_transforms(s12, r12, t12)
    on, om12 = oracle.search_by_sim3(oracle_frame(kf1, TUM), oracle_frame(kf2, TUM), SF, cam, poses["t1w"], poses["t2w"], T21, T12, pts1, pts2, th)
    rn, rm12 = refm.search_by_sim3(ref_frame(kf1, TUM, poses["t1w"]), ref_frame(kf2, TUM, poses["t2w"]), pts1, pts2, s12, r12, t12, th)
    assert on == rn and np.array_equal(om12, rm12) and on > 20


def test_distinctive_descriptors():
    for seed in (1, 2, 3):
        dd, ds = synth.observation_descriptors(300, seed)
        ob = oracle.distinctive_descriptors(dd, ds)
        rb = refm.distinctive_descriptors(dd, ds)
        dd = np.asarray(dd).reshape(-1, 32)
        for p in range(len(ds) - 1):
            if ds[p + 1] == ds[p]:
                assert ob[p] == -1 and rb[p] == -1
            else:
                assert np.array_equal(dd[ds[p] + ob[p]], dd[ds[p] + rb[p]]), p
