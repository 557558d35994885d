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
