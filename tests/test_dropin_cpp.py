"""The C++ drop-in (object_slam_b200/host/ORBextractor.{h,cc}: same class surface as the reference's
include/ORBextractor.h) compiles against a cv:: API and, on the GPU box, reproduces the oracle."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle
from object_slam_b200 import synth, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "object_slam_b200", "host")


def _build(out):
    # OpenCV's C++ headers are not in this image; oracle/cvshim models the slice of cv:: the reference's
    # extractor (and therefore this drop-in) uses.  In the reference tree the same two files compile
    # against the real OpenCV.
    cmd = ["g++", "-std=gnu++11", "-O2", "-I" + os.path.join(ROOT, "oracle", "cvshim"), "-I" + os.path.join(ROOT, "include"),
           "-I" + HOST, "-o", out, os.path.join(HOST, "dropin_check.cpp"), os.path.join(HOST, "ORBextractor.cc"),
           "-L" + os.path.dirname(_capi.LIB_PATH), "-lobslam_b200", "-Wl,-rpath," + os.path.dirname(_capi.LIB_PATH)]
    subprocess.check_call(cmd)


def test_dropin_compiles_and_links():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "dropin_check")
        _build(exe)
        assert os.path.exists(exe)


def test_dropin_class_surface_matches_reference_header():
    ref = "/root/reference/include/ORBextractor.h"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box")
    import re
    mine = open(os.path.join(HOST, "ORBextractor.h")).read()
    theirs = open(ref).read()
    for name in ("GetLevels", "GetScaleFactor", "GetScaleFactors", "GetInverseScaleFactors", "GetScaleSigmaSquares",
                 "GetInverseScaleSigmaSquares", "mvImagePyramid", "operator()"):
        assert name in theirs and name in mine
    ctor = re.search(r"ORBextractor\(int nfeatures, float scaleFactor, int nlevels,\s*int iniThFAST, int minThFAST\)", theirs)
    assert ctor and re.search(r"ORBextractor\(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST\)", mine)


@pytest.mark.gpu
def test_dropin_matches_oracle(gpu):
    shape = synth.KITTI_SHAPE
    L, R = synth.stereo_pair(shape, 9)
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "dropin_check")
        _build(exe)
        L.tofile(os.path.join(d, "l.raw")); R.tofile(os.path.join(d, "r.raw"))
        out = os.path.join(d, "out")
        subprocess.check_call([exe, str(shape[1]), str(shape[0]), "2000", os.path.join(d, "l.raw"), out,
                               os.path.join(d, "r.raw"), str(synth.KITTI_BF), str(synth.KITTI_FX)])
        k = np.fromfile(out + ".kp", dtype=oracle.KEYPOINT_DTYPE)
        desc = np.fromfile(out + ".desc", dtype=np.uint8).reshape(-1, 32)
        ur = np.fromfile(out + ".uright", dtype=np.float32)
        dp = np.fromfile(out + ".depth", dtype=np.float32)
    oL, oR = oracle.OracleExtractor(2000), oracle.OracleExtractor(2000)
    kL, dL = oL(L); kR, dR = oR(R)
    t = oL.tables()
    our, odp, _ = oracle.stereo_match(kL, dL, kR, dR, [oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)],
                                      t["scale"], t["inv_scale"], synth.KITTI_BF, 0.0, synth.KITTI_FX)
    assert k.tobytes() == kL.tobytes() and np.array_equal(desc, dL)
    assert np.array_equal(ur, our) and np.array_equal(dp, odp)


# ----------------------------------------------------------------------------- the matcher mirror (host/ORBmatcher.h)
def _build_matcher(out):
    cmd = ["g++", "-std=gnu++11", "-O2", "-I" + os.path.join(ROOT, "oracle", "cvshim"), "-I" + os.path.join(ROOT, "include"),
           "-I" + HOST, "-o", out, os.path.join(HOST, "matcher_check.cpp"),
           "-L" + os.path.dirname(_capi.LIB_PATH), "-lobslam_b200", "-Wl,-rpath," + os.path.dirname(_capi.LIB_PATH)]
    subprocess.check_call(cmd)


def _frame_bytes(frame, shape):
    k, d, ur = frame
    cam = synth.camera_for(shape)
    hdr = np.array([len(k), 0 if ur is None else 1], np.int32).tobytes()
    fl = np.array([0, shape[1], 0, shape[0], *cam], np.float32).tobytes()
    out = hdr + fl + synth.scale_factors().tobytes() + np.ascontiguousarray(k).tobytes() + np.ascontiguousarray(d).tobytes()
    if ur is not None:
        out += np.ascontiguousarray(ur, np.float32).tobytes()
    return out


def test_matcher_mirror_compiles_and_names_match_reference_header():
    with tempfile.TemporaryDirectory() as d:
        _build_matcher(os.path.join(d, "matcher_check"))
    ref = "/root/reference/include/ORBmatcher.h"
    if os.path.exists(ref):
        theirs, mine = open(ref).read(), open(os.path.join(HOST, "ORBmatcher.h")).read()
        for name in ("ORBmatcher(float nnratio=0.6, bool checkOri=true)", "static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b)",
                     "int SearchByProjection(", "int SearchForInitialization(", "int SearchByBoW(", "TH_LOW", "TH_HIGH", "HISTO_LENGTH", "ComputeThreeMaxima",
                     "mfNNratio", "mbCheckOrientation"):
            assert name in theirs and name in mine, name


@pytest.mark.gpu
def test_matcher_mirror_matches_oracle(gpu):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import matcher_cases as mc
    shape = synth.TUM_SHAPE
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "matcher_check")
        _build_matcher(exe)
        # SearchByProjection(Frame&, const vector<MapPoint*>&, th)
        frame, mp, _ = mc.map_case(shape, 1000, 5000, 11)
        blob = _frame_bytes(frame, shape) + np.array([5000], np.int32).tobytes() + np.array([3.0, 0.8], np.float32).tobytes()
        for key, dt in zip(mc.MP_KEYS, (np.uint8, np.float32, np.float32, np.float32, np.int32, np.float32, np.uint8, np.int32)):
            blob += np.ascontiguousarray(mp[key], dt).tobytes()
        open(os.path.join(d, "map.bin"), "wb").write(blob)
        subprocess.check_call([exe, "map", os.path.join(d, "map.bin"), os.path.join(d, "map.out")])
        res = np.fromfile(os.path.join(d, "map.out"), np.int32)
        on, om = mc.oracle_map(frame, shape, mp, 3.0, 0.8)
        assert res[0] == on and np.array_equal(res[1:], om)
        # SearchForInitialization
        f1, f2, prev = synth.init_pair(shape, 1000, 12)
        blob = _frame_bytes(f1, shape) + _frame_bytes(f2, shape) + np.array([100], np.int32).tobytes() + np.array([0.9], np.float32).tobytes()
        blob += np.ascontiguousarray(prev, np.float32).tobytes()
        open(os.path.join(d, "init.bin"), "wb").write(blob)
        subprocess.check_call([exe, "init", os.path.join(d, "init.bin"), os.path.join(d, "init.out")])
        raw = open(os.path.join(d, "init.out"), "rb").read()
        n = np.frombuffer(raw[:4], np.int32)[0]
        m12 = np.frombuffer(raw[4:4 + 4000], np.int32)
        pm = np.frombuffer(raw[4 + 4000:], np.float32).reshape(-1, 2)
        on, om12, opm = mc.oracle_init(f1, f2, shape, prev, 100, 0.9)
        assert n == on and np.array_equal(m12, om12) and np.array_equal(pm, opm)
        # SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) through std::map feature vectors
        import oracle
        a, b, _ = synth.bow_pair(shape, 1000, 13, n_nodes=80)
        blob = np.array([0.7], np.float32).tobytes()
        for sd in (a, b):
            node = np.zeros(sd["n"], np.int32)
            for k in range(len(sd["node_id"])):
                node[sd["node_idx"][sd["node_start"][k]:sd["node_start"][k + 1]]] = sd["node_id"][k]
            blob += np.array([sd["n"]], np.int32).tobytes() + np.ascontiguousarray(sd["keys_un"]).tobytes() + \
                np.ascontiguousarray(sd["descriptors"]).tobytes() + np.ascontiguousarray(sd["valid"], np.uint8).tobytes() + node.tobytes()
        open(os.path.join(d, "bow.bin"), "wb").write(blob)
        subprocess.check_call([exe, "bow", os.path.join(d, "bow.bin"), os.path.join(d, "bow.out")])
        res = np.fromfile(os.path.join(d, "bow.out"), np.int32)
        on, _, om21 = oracle.search_by_bow(a, dict(b, valid=None), 50, False, 0.7, True)
        assert res[0] == on and np.array_equal(res[1:], om21)
