"""Oracle == the reference itself.  The committed fixtures in tests/golden/ were produced by the
reference's own, unmodified src/ORBextractor.cc (compiled in place into oracle/_ref, see
tools/make_golden.py); where oracle/_ref is present the two are also compared live."""
import glob
import os

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _level_checksums(levels):
    return np.array([[int(p.sum()), int((p.astype(np.int64) * (np.arange(p.size).reshape(p.shape) % 65521)).sum())]
                     for p in levels], np.int64)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "extract_*.npz"))), ids=os.path.basename)
def test_oracle_matches_golden(path):
    g = np.load(path)
    img = getattr(synth, str(g["generator"]))(tuple(int(v) for v in g["shape"]), int(g["seed"]))
    e = oracle.OracleExtractor(int(g["nfeatures"]))
    k, d = e(img)
    assert k.tobytes() == g["keypoints"].tobytes()
    assert np.array_equal(d, g["descriptors"])
    assert np.array_equal(_level_checksums([e.level(l) for l in range(8)]), g["level_checksums"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "stereo_*.npz"))), ids=os.path.basename)
def test_stereo_oracle_matches_golden(path):
    g = np.load(path)
    L, R = synth.stereo_pair(synth.KITTI_SHAPE, int(g["seed"]))
    oL, oR = oracle.OracleExtractor(2000), oracle.OracleExtractor(2000)
    kL, dL = oL(L); kR, dR = oR(R)
    t = oL.tables()
    ur, dp, sad = oracle.stereo_match(kL, dL, kR, dR, [oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)],
                                      t["scale"], t["inv_scale"], synth.KITTI_BF, 0.0, synth.KITTI_FX)
    assert len(kL) == int(g["n_left"]) and len(kR) == int(g["n_right"])
    assert np.array_equal(ur, g["uRight"]) and np.array_equal(dp, g["depth"]) and np.array_equal(sad, g["sad"])


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape,nf", [(synth.TUM_SHAPE, 1000), (synth.KITTI_SHAPE, 2000), ((240, 320), 500)])
@pytest.mark.parametrize("gen", ["blocky_image", "noise_image"])
def test_oracle_matches_compiled_reference(shape, nf, gen):
    for seed in (11, 12):
        img = getattr(synth, gen)(shape, seed)
        k, d = oracle.OracleExtractor(nf)(img)
        ref = oracle.ReferenceExtractor(nf)
        rk, rd = ref(img)
        assert k.tobytes() == rk.tobytes()
        assert np.array_equal(d, rd)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_tables_match_compiled_reference():
    for nf, sf, nl in ((1000, 1.2, 8), (2000, 1.2, 8), (1500, 1.1, 6)):
        ref = oracle.ReferenceExtractor(nf, sf, nl)
        sc = np.empty(nl, np.float32)
        oracle.ref_lib().ref_get_scale_factors(ref._h, sc.ctypes.data_as(oracle.C.c_void_p))
        t = oracle.OracleExtractor(nf, sf, nl).tables()
        assert np.array_equal(sc, t["scale"])
        assert int(t["features_per_level"].sum()) == nf


def test_oracle_edge_cases():
    e = oracle.OracleExtractor(1000)
    k, d = e(synth.flat_image((480, 640)))
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = e(synth.blocky_image((70, 90), 0))       # upper levels smaller than one cell
    assert len(k) >= 0
    umax = e.tables()["umax"]
    assert list(umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    for i in range(50):
        assert oracle.descriptor_distance(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())
    assert oracle.descriptor_distance(a[0], a[0]) == 0
    assert oracle.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256
