"""Pins the oracle's restated OpenCV primitives to cv2 4.13 (the build this image ships) and its
sincosf model to this image's libm -- the third-party arithmetic the reference's ORBextractor rests on
(OpenCV is an un-vendored, un-pinned dependency of the reference: CMakeLists.txt:31-37)."""
import os

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth

cv2 = pytest.importorskip("cv2")


def _images():
    rng = np.random.default_rng(7)
    yield "noise", rng.integers(0, 256, (97, 131), dtype=np.uint8)
    yield "blocky", synth.blocky_image((120, 160), 3)
    yield "binary", (rng.integers(0, 2, (64, 75)) * 255).astype(np.uint8)
    g = np.linspace(0, 255, 201)[None, :] * np.ones((53, 1))
    yield "ramp", g.astype(np.uint8)


@pytest.mark.parametrize("name,img", list(_images()))
def test_resize_matches_cv2(name, img):
    h, w = img.shape
    for s in (1.2, 1.5, 2.0, 1.07):
        dw, dh = int(round(w / s)), int(round(h / s))
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.resize_linear(img, dw, dh), ref), (name, s)


def test_resize_pyramid_shapes_match_cv2():
    for shape in (synth.TUM_SHAPE, synth.KITTI_SHAPE):
        img = synth.blocky_image(shape, 1)
        e = oracle.OracleExtractor(1000)
        e(img)
        prev = img
        for l in range(1, 8):
            cur = e.level(l)
            assert np.array_equal(cv2.resize(prev, (cur.shape[1], cur.shape[0]), interpolation=cv2.INTER_LINEAR), cur)
            prev = cur


@pytest.mark.parametrize("name,img", list(_images()))
def test_gaussian_matches_cv2(name, img):
    ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(oracle.gaussian7x7(img), ref)


@pytest.mark.parametrize("name,img", list(_images()))
@pytest.mark.parametrize("th", [7, 20, 40])
def test_fast_matches_cv2(name, img, th):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in det.detect(img)], np.int32).reshape(-1, 3)
    got = oracle.fast9_16(img, th, nms=True)
    assert np.array_equal(got, ref)
    assert np.array_equal(oracle.fast9_16(img, th, nms=True, simple=True), ref)


def test_fast_score_map_consistent():
    img = synth.noise_image((60, 70), 2)
    sm = oracle.fast_score_map(img)
    for th in (7, 20):
        k = oracle.fast9_16(img, th, nms=False)
        mask = np.zeros_like(sm, bool)
        mask[k[:, 1], k[:, 0]] = True
        assert np.array_equal(mask, sm >= th)       # a pixel is a corner at t iff score >= t


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(3)
    y = rng.integers(-200000, 200000, 200000).astype(np.float32)
    x = rng.integers(-200000, 200000, 200000).astype(np.float32)
    y[:10] = 0; x[5:15] = 0
    # the scalar cv::fastAtan2 the reference calls (ORBextractor.cc:103); cv2.phase's SIMD path rounds differently
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(oracle.fast_atan2(y, x), ref)


def test_sincosf_model_matches_libm():
    # the device evaluates glibc's sincosf algorithm in binary64; the model must equal libm bit for bit
    ang = (np.arange(0, 360000, dtype=np.float32) * np.float32(0.001))
    rad = ang * np.float32(np.float32(np.pi) / np.float32(180.0))
    s_m, c_m = oracle.sincosf(rad, model=True)
    s_l, c_l = oracle.sincosf(rad, model=False)
    assert np.array_equal(s_m, s_l) and np.array_equal(c_m, c_l)
    # strided sweep over every binary32 in [1e-7, 2*pi]
    assert oracle.lib().orc_sincosf_sweep(1e-7, 6.2831855, 97) == 0


def test_pattern_matches_reference_file():
    path = "/root/reference/src/ORBextractor.cc"
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    import re
    src = open(path).read()
    body = src[src.index("static int bit_pattern_31_[256*4]"):]
    body = body[body.index("{") + 1:body.index("};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    vals = np.array([int(v) for v in re.findall(r"-?\d+", body)], np.int32).reshape(256, 4)
    assert np.array_equal(vals, oracle.pattern())
