"""Where does the end-to-end step time go?  (run on the GPU box)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from object_slam_b200 import synth
from object_slam_b200._capi import pinned_empty, KEYPOINT_DTYPE
from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches
F, (H, W) = 64, synth.KITTI_SHAPE
pairs = [synth.stereo_pair((H, W), i) for i in range(8)]
exL = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=F); exR = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=F)
cap = exL.capacity
pinL = pinned_empty((F, H, W), np.uint8); pinR = pinned_empty((F, H, W), np.uint8)
for i in range(F): pinL[i] = pairs[i % 8][0]; pinR[i] = pairs[i % 8][1]
mk = lambda: (pinned_empty((F, cap), KEYPOINT_DTYPE), pinned_empty((F, cap, 32), np.uint8), pinned_empty((F,), np.int32))
outL, outR = mk(), mk()
outS = (pinned_empty((F, cap), np.float32), pinned_empty((F, cap), np.float32))
def t(fn, n=10):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return 1e3 * (time.perf_counter() - t0) / n
print("one eye extract_batch (pinned in/out): %.2f ms" % t(lambda: exL.extract_batch(pinL, out=outL, copy=False)))
def both():
    th = threading.Thread(target=lambda: exR.extract_batch(pinR, out=outR, copy=False)); th.start()
    exL.extract_batch(pinL, out=outL, copy=False); th.join()
print("two eyes, two threads: %.2f ms" % t(both))
print("stereo match (host out): %.2f ms" % t(lambda: ComputeStereoMatches(exL, exR, synth.KITTI_BF, 0.0, synth.KITTI_FX, out=outS)))
import torch
x = torch.from_numpy(pinL); d = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
def h2d(): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
print("raw H2D of one eye batch (%.1f MB): %.2f ms" % (pinL.nbytes / 1e6, t(h2d)))
y = torch.from_numpy(outL[1]); dd = torch.empty(y.shape, dtype=torch.uint8, device="cuda")
def d2h(): y.copy_(dd, non_blocking=True); torch.cuda.synchronize()
print("raw D2H of descriptors (%.1f MB): %.2f ms" % (outL[1].nbytes / 1e6, t(d2h)))
