#!/usr/bin/env python
"""Single-frame latency of the host-buffer API (the real-time use of the front end: one stereo frame at a time)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from object_slam_b200 import synth
from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches
L, R = synth.stereo_pair(synth.KITTI_SHAPE, 1)
eL = ORBextractor(2000, 1.2, 8, 20, 7, max_size=(1241, 376)); eR = ORBextractor(2000, 1.2, 8, 20, 7, max_size=(1241, 376))
def mono():
    return eL(L)
def stereo():
    th = threading.Thread(target=lambda: eR(R)); th.start(); eL(L); th.join()
    return ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
for name, fn in (("mono extract 1241x376", mono), ("stereo extract + match", stereo)):
    for _ in range(20): fn()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(f"{name}: median {np.median(ts):.3f} ms, p90 {np.percentile(ts, 90):.3f} ms, min {ts.min():.3f} ms")
