"""Parity of the CUDA Frame::ComputeStereoMatches with the oracle: match set, uRight and depth.
The contract for the sub-pixel disparity is 1e-3 px; the float path is restated op for op, so the
values are compared bit for bit and the tolerance is asserted on top."""
import glob
import os

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pair(shape, nf, **kw):
    from object_slam_b200.extractor import ORBextractor
    return (ORBextractor(nf, 1.2, 8, 20, 7, max_size=(shape[1], shape[0]), **kw),
            ORBextractor(nf, 1.2, 8, 20, 7, max_size=(shape[1], shape[0]), **kw))


def _oracle_stereo(L, R, nf, mbf, minD, maxD):
    oL, oR = oracle.OracleExtractor(nf), oracle.OracleExtractor(nf)
    kL, dL = oL(L); kR, dR = oR(R)
    t = oL.tables()
    return oracle.stereo_match(kL, dL, kR, dR, [oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)],
                               t["scale"], t["inv_scale"], mbf, minD, maxD)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "stereo_*.npz"))), ids=os.path.basename)
def test_matches_golden(gpu, path):
    from object_slam_b200.extractor import ComputeStereoMatches
    g = np.load(path)
    L, R = synth.stereo_pair(synth.KITTI_SHAPE, int(g["seed"]))
    eL, eR = _pair(synth.KITTI_SHAPE, 2000)
    eL(L); eR(R)
    (ur, dp), = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
    assert np.array_equal(ur >= 0, g["uRight"] >= 0)                 # same match set
    assert np.abs(ur - g["uRight"]).max() <= 1e-3                     # contract
    assert np.array_equal(ur, g["uRight"]) and np.array_equal(dp, g["depth"])


@pytest.mark.parametrize("shape,nf,mbf,maxD", [(synth.KITTI_SHAPE, 2000, synth.KITTI_BF, synth.KITTI_FX),
                                               (synth.TUM_SHAPE, 1000, 40.0, 525.0)])
def test_batch_matches_oracle(gpu, shape, nf, mbf, maxD):
    from object_slam_b200.extractor import ComputeStereoMatches
    pairs = [synth.stereo_pair(shape, s) for s in range(30, 36)]
    eL, eR = _pair(shape, nf, max_batch=6)
    eL.extract_batch([p[0] for p in pairs]); eR.extract_batch([p[1] for p in pairs])
    res = ComputeStereoMatches(eL, eR, mbf, 0.0, maxD)
    total = 0
    for (L, R), (ur, dp) in zip(pairs, res):
        our, odp, _ = _oracle_stereo(L, R, nf, mbf, 0.0, maxD)
        assert np.array_equal(ur, our) and np.array_equal(dp, odp)
        total += int((ur >= 0).sum())
        ok = ur >= 0
        assert np.all(dp[ok] > 0)
    assert total > 100


def test_disparity_limits_and_empty(gpu):
    from object_slam_b200.extractor import ComputeStereoMatches
    shape = synth.KITTI_SHAPE
    L, R = synth.stereo_pair(shape, 3)
    eL, eR = _pair(shape, 2000)
    eL(L); eR(R)
    for minD, maxD in ((0.0, 30.0), (10.0, 60.0), (0.0, 1e-3)):
        (ur, dp), = ComputeStereoMatches(eL, eR, synth.KITTI_BF, minD, maxD)
        our, odp, _ = _oracle_stereo(L, R, 2000, synth.KITTI_BF, minD, maxD)
        assert np.array_equal(ur, our) and np.array_equal(dp, odp)
    # unrelated right image: (almost) nothing matches, the empty accepted list must not crash
    eR(synth.noise_image(shape, 1))
    (ur, dp), = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
    our, odp, _ = _oracle_stereo(L, synth.noise_image(shape, 1), 2000, synth.KITTI_BF, 0.0, synth.KITTI_FX)
    assert np.array_equal(ur, our) and np.array_equal(dp, odp)
    # textureless left image: zero keypoints
    eL(synth.flat_image(shape))
    (ur, dp), = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
    assert len(ur) == 0


def test_identical_eyes_give_zero_disparity_clamp(gpu):
    """Left == right: disparity <= 0 takes the 0.01 clamp branch (Frame.cc:852-856)."""
    from object_slam_b200.extractor import ComputeStereoMatches
    shape = synth.TUM_SHAPE
    img = synth.blocky_image(shape, 4)
    eL, eR = _pair(shape, 1000)
    eL(img); eR(img)
    (ur, dp), = ComputeStereoMatches(eL, eR, 40.0, 0.0, 525.0)
    our, odp, _ = _oracle_stereo(img, img, 1000, 40.0, 0.0, 525.0)
    assert np.array_equal(ur, our) and np.array_equal(dp, odp)
    # every accepted match has SAD 0 here, so the median cut (thDist = 0) removes them all -- as in the reference
