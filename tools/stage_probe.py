#!/usr/bin/env python
"""Per-stage CUDA-event times of ONE KITTI image (the live-frame case): where the 0.15 ms of a single extraction go."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from object_slam_b200 import synth
from object_slam_b200.extractor import ORBextractor
H, W = synth.KITTI_SHAPE
L, R = synth.stereo_pair(synth.KITTI_SHAPE, 1)
e = ORBextractor(2000, 1.2, 8, 20, 7, max_size=(W, H))
for _ in range(20): e(L)
e.set_profiling(True)
for _ in range(50): e(L)
ms, n, ns = e.stage_ms()
print({k: round(1e3 * v / max(n, 1), 1) for k, v in ms.items()}, "us per call over", n, "calls")
