"""Parity of the CUDA extractor (through the C ABI) with the oracle and the reference-generated
golden fixtures.  Bit-exact for pyramid pixels, FAST candidates, selected keypoints and their order,
blurred levels, descriptors and every KeyPoint field (the angle contract is 1e-3 deg; it holds at 0)."""
import glob
import os

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth
from object_slam_b200 import _capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ex(nf, shape, **kw):
    from object_slam_b200.extractor import ORBextractor
    return ORBextractor(nf, 1.2, 8, 20, 7, max_size=(shape[1], shape[0]), **kw)


def _assert_same(k, d, ok, od):
    assert len(k) == len(ok)
    for f in ok.dtype.names:
        if f == "angle":
            assert np.abs(k[f] - ok[f]).max(initial=0) <= 1e-3       # contract; observed 0
        assert np.array_equal(k[f], ok[f]), f
    assert np.array_equal(d, od)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "extract_*.npz"))), ids=os.path.basename)
def test_matches_reference_golden(gpu, path):
    g = np.load(path)
    shape = tuple(int(v) for v in g["shape"])
    img = getattr(synth, str(g["generator"]))(shape, int(g["seed"]))
    k, d = _ex(int(g["nfeatures"]), shape)(img)
    _assert_same(k, d, g["keypoints"], g["descriptors"])


@pytest.mark.parametrize("shape,nf", [(synth.TUM_SHAPE, 1000), (synth.KITTI_SHAPE, 2000)])
@pytest.mark.parametrize("gen", ["blocky_image", "noise_image"])
def test_every_stage_matches_oracle(gpu, shape, nf, gen):
    img = getattr(synth, gen)(shape, 21)
    o = oracle.OracleExtractor(nf)
    ok, od = o(img)
    e = _ex(nf, shape)
    k, d = e(img)
    for l in range(8):
        assert np.array_equal(e.level(l), o.level(l)), f"pyramid level {l}"
        b = o.level(l, True)
        if b is not None:
            assert np.array_equal(e.level(l, blurred=True), b), f"blurred level {l}"
        oc = o.level_keypoints(l, False)
        assert np.array_equal(e.candidates(l), np.stack([oc["x"], oc["y"], oc["response"]], 1).astype(np.int32)), f"candidates {l}"
        os_ = o.level_keypoints(l, True)
        assert np.array_equal(e.selected(l), np.stack([os_["x"] - 16, os_["y"] - 16, os_["response"]], 1).astype(np.int32)), f"selected {l}"
    _assert_same(k, d, ok, od)


def test_many_seeds_end_to_end(gpu):
    """>= 100 frames over two shapes and two generators, mismatch counts must be zero."""
    bad = 0
    for shape, nf in ((synth.TUM_SHAPE, 1000), (synth.KITTI_SHAPE, 2000)):
        e = _ex(nf, shape, max_batch=13)
        o = oracle.OracleExtractor(nf)
        for gen, seeds in (("blocky_image", range(100, 126)), ("noise_image", range(200, 226))):
            imgs = [getattr(synth, gen)(shape, s) for s in seeds]
            for i in range(0, len(imgs), 13):
                res = e.extract_batch(imgs[i:i + 13])
                for im, (k, d) in zip(imgs[i:i + 13], res):
                    ok, od = o(im)
                    bad += int(k.tobytes() != ok.tobytes()) + int(not np.array_equal(d, od))
    assert bad == 0


def test_batch_and_device_api_equal_single(gpu):
    import torch
    shape = synth.KITTI_SHAPE
    imgs = [synth.blocky_image(shape, s) for s in range(40, 45)]
    single = _ex(2000, shape)
    want = [single(im) for im in imgs]
    eb = _ex(2000, shape, max_batch=5)
    for (k, d), (wk, wd) in zip(eb.extract_batch(imgs), want):
        assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd)
    # device-resident images with an arbitrary (16-byte aligned) pitch, on torch's current stream
    pitch = 1264
    host = np.zeros((5, shape[0], pitch), np.uint8)
    for i, im in enumerate(imgs):
        host[i, :, :shape[1]] = im
    dev = torch.from_numpy(host).cuda()
    ts = torch.cuda.Stream()
    ts.wait_stream(torch.cuda.current_stream())
    eb.extract_device(dev.data_ptr(), 5, shape[1], shape[0], pitch, shape[0] * pitch, ts.cuda_stream)
    for (k, d), (wk, wd) in zip(eb.fetch(), want):
        assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd)
    assert np.array_equal(eb.level(0, image_index=3), imgs[3])


def test_shape_changes_between_calls(gpu):
    e = _ex(1000, (480, 1241))
    for shape in ((240, 320), synth.TUM_SHAPE, (376, 1241), (100, 300), (240, 320)):
        img = synth.blocky_image(shape, 3)
        k, d = e(img)
        ok, od = oracle.OracleExtractor(1000)(img)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), shape


def test_edge_cases(gpu):
    e = _ex(1000, synth.TUM_SHAPE)
    # textureless: zero keypoints, descriptors released (0 x 32)
    k, d = e(synth.flat_image(synth.TUM_SHAPE))
    assert len(k) == 0 and d.shape == (0, 32)
    # empty image: outputs untouched / empty (ORBextractor.cc:1046)
    k, d = e(np.empty((0, 0), np.uint8))
    assert len(k) == 0
    # upper pyramid levels smaller than one FAST cell
    for shape in ((70, 90), (64, 64), (45, 200)):
        img = synth.noise_image(shape, 5)
        k, d = e(img)
        ok, od = oracle.OracleExtractor(1000)(img)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), shape
    # every cell falls back to minThFAST: low-contrast texture (|diff| < 20 everywhere)
    rng = np.random.default_rng(9)
    low = (120 + rng.integers(0, 17, synth.TUM_SHAPE)).astype(np.uint8)
    k, d = e(low)
    ok, od = oracle.OracleExtractor(1000)(low)
    assert len(ok) > 0 and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # N not reached: very few corners
    sparse = synth.flat_image(synth.TUM_SHAPE)
    sparse[100:140, 200:260] = 255
    sparse[300:320, 400:410] = 0
    k, d = e(sparse)
    ok, od = oracle.OracleExtractor(1000)(sparse)
    assert 0 < len(ok) < 200 and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # non-contiguous rows (stride > width)
    big = synth.blocky_image((480, 700), 8)
    view = big[:, :640]
    k, d = e(view)
    ok, od = oracle.OracleExtractor(1000)(np.ascontiguousarray(view))
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


def test_other_parameters(gpu):
    from object_slam_b200.extractor import ORBextractor
    img = synth.blocky_image(synth.TUM_SHAPE, 17)
    for nf, sf, nl, ini, mn in ((2000, 1.2, 8, 20, 7), (500, 1.1, 6, 12, 5), (300, 1.5, 4, 30, 10), (4000, 1.2, 8, 20, 7)):
        e = ORBextractor(nf, sf, nl, ini, mn, max_size=(640, 480))
        k, d = e(img)
        ok, od = oracle.OracleExtractor(nf, sf, nl, ini, mn)(img)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), (nf, sf, nl, ini, mn)
        t = oracle.OracleExtractor(nf, sf, nl, ini, mn).tables()
        assert np.array_equal(e.GetScaleFactors(), t["scale"]) and np.array_equal(e.GetInverseScaleFactors(), t["inv_scale"])
        assert np.array_equal(e.GetScaleSigmaSquares(), t["sigma2"]) and np.array_equal(e.GetInverseScaleSigmaSquares(), t["inv_sigma2"])
        assert np.array_equal(e.mnFeaturesPerLevel, t["features_per_level"])


def test_error_codes(gpu):
    from object_slam_b200.extractor import ORBextractor
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_size=(640, 480))
    with pytest.raises(_capi.ObsError) as ei:
        e(synth.blocky_image((600, 800), 0))            # larger than max_size
    assert ei.value.code == _capi.OBS_ERR_INVALID
    with pytest.raises(_capi.ObsError) as ei:
        e.extract_batch([synth.blocky_image(synth.TUM_SHAPE, 0)] * 2)     # exceeds max_batch
    assert ei.value.code == _capi.OBS_ERR_CAPACITY
    with pytest.raises(ValueError):
        e(np.zeros((480, 640, 3), np.uint8))


def test_concurrent_handles_two_threads(gpu):
    """Two extractors driven from two host threads, as Frame::Frame does for the stereo eyes (Frame.cc:78-81)."""
    import threading
    shape = synth.KITTI_SHAPE
    L, R = synth.stereo_pair(shape, 77)
    eL, eR = _ex(2000, shape), _ex(2000, shape)
    out = {}
    for rep in range(5):
        tl = threading.Thread(target=lambda: out.__setitem__("L", eL(L)))
        tr = threading.Thread(target=lambda: out.__setitem__("R", eR(R)))
        tl.start(); tr.start(); tl.join(); tr.join()
        for key, im in (("L", L), ("R", R)):
            ok, od = oracle.OracleExtractor(2000)(im)
            assert out[key][0].tobytes() == ok.tobytes() and np.array_equal(out[key][1], od)


def test_pinned_host_buffers_take_the_dma_path(gpu):
    """Page-locked inputs/outputs (obs_host_alloc) must give the same bytes as pageable ones."""
    from object_slam_b200._capi import pinned_empty, KEYPOINT_DTYPE
    from object_slam_b200.extractor import ComputeStereoMatches
    shape = synth.KITTI_SHAPE
    pairs = [synth.stereo_pair(shape, s) for s in (60, 61, 62)]
    eL, eR = _ex(2000, shape, max_batch=3), _ex(2000, shape, max_batch=3)
    want = eL.extract_batch([p[0] for p in pairs]); eR.extract_batch([p[1] for p in pairs])
    wantS = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
    cap = eL.capacity
    pinL, pinR = pinned_empty((3,) + shape, np.uint8), pinned_empty((3,) + shape, np.uint8)
    for i, (l, r) in enumerate(pairs):
        pinL[i] = l; pinR[i] = r
    out = (pinned_empty((3, cap), KEYPOINT_DTYPE), pinned_empty((3, cap, 32), np.uint8), pinned_empty((3,), np.int32))
    outR = (pinned_empty((3, cap), KEYPOINT_DTYPE), pinned_empty((3, cap, 32), np.uint8), pinned_empty((3,), np.int32))
    got = eL.extract_batch(pinL, out=out, copy=False); eR.extract_batch(pinR, out=outR, copy=False)
    outS = (pinned_empty((3, cap), np.float32), pinned_empty((3, cap), np.float32))
    gotS = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX, out=outS)
    for (k, d), (wk, wd) in zip(got, want):
        assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd)
    for (u, z), (wu, wz) in zip(gotS, wantS):
        assert np.array_equal(u, wu) and np.array_equal(z, wz)


def test_very_wide_image_capacity(gpu):
    """A level much wider than high starts DistributeOctTree with many roots and its first sweep always splits them all, so a level
    can end far above its quota (src/ORBextractor.cc:606-669): 402 keypoints for nFeatures = 300 here.  The capacity must follow."""
    from object_slam_b200.extractor import ORBextractor
    shape = (118, 1121)
    img = synth.blocky_image(shape, 651119381)
    e = ORBextractor(300, 1.1, 8, 20, 5, max_size=(shape[1], shape[0]))
    k, d = e(img)
    ok, od = oracle.OracleExtractor(300, 1.1, 8, 20, 5)(img)
    assert len(ok) > 300 + 4 * 8
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


def test_alternating_handles_with_different_settings(gpu):
    """Handles with different nfeatures / shapes used alternately (the monocular drop-in: mpIniORBextractor at 2 x nFeatures beside
    mpORBextractorLeft): the dynamic shared-memory limit of a kernel is process-wide, so a handle with smaller node or tile
    capacities must not lower it under the other one's feet."""
    big = _ex(2000, synth.KITTI_SHAPE)
    small = _ex(500, synth.TUM_SHAPE)
    huge = _ex(4000, synth.KITTI_SHAPE)
    imgK = synth.blocky_image(synth.KITTI_SHAPE, 31)
    imgT = synth.blocky_image(synth.TUM_SHAPE, 32)
    want = {}
    for name, e, img, nf in (("big", big, imgK, 2000), ("small", small, imgT, 500), ("huge", huge, imgK, 4000)):
        want[name] = oracle.OracleExtractor(nf)(img)
    for _ in range(3):                      # geometry unchanged between rounds: nothing is re-prepared
        for name, e, img in (("big", big, imgK), ("small", small, imgT), ("huge", huge, imgK), ("small", small, imgT), ("big", big, imgK)):
            k, d = e(img)
            _assert_same(k, d, *want[name])
