#!/usr/bin/env python
"""Single-frame latency of the host-buffer API (the real-time use of the front end: one stereo frame at a time).
usage: latency_probe.py [out.json]"""
import json, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from object_slam_b200 import synth
from object_slam_b200._capi import lib, check
from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches, StereoFrames
H, W = synth.KITTI_SHAPE
L, R = synth.stereo_pair(synth.KITTI_SHAPE, 1)
eL = ORBextractor(2000, 1.2, 8, 20, 7, max_size=(W, H)); eR = ORBextractor(2000, 1.2, 8, 20, 7, max_size=(W, H))
pipe = StereoFrames(2000, 1.2, 8, 20, 7, (W, H), 1)
pipe.left[0] = L; pipe.right[0] = R
def mono():
    return eL(L)
import ctypes as C
from object_slam_b200._capi import KEYPOINT_DTYPE, ptr
_cap = eL.capacity
_kps = np.empty(_cap, KEYPOINT_DTYPE); _desc = np.empty((_cap, 32), np.uint8); _n = C.c_int(0)
_kps[:] = 0; _desc[:] = 0                                   # touch the pages once
def mono_c():                                               # the C call alone, caller-owned (pageable) arrays reused: what the C++ drop-in does
    check(lib().obs_extract(eL._h, ptr(L), W, H, W, ptr(_kps), ptr(_desc), _cap, C.byref(_n)))
def stereo_threads():
    th = threading.Thread(target=lambda: eR(R)); th.start(); eL(L); th.join()
    return ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX)
def stereo_one_call():
    pipe.submit(synth.KITTI_BF, 0.0, synth.KITTI_FX); pipe.wait()
out = {}
def run(name, fn, opts):
    for k, v in opts.items(): check(lib().obs_set_option(k.encode(), v))
    for _ in range(30): fn()
    ts = []
    for _ in range(300):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    out[name] = {"median_ms": float(np.median(ts)), "p90_ms": float(np.percentile(ts, 90)), "min_ms": float(ts.min())}
    print(f"{name}: median {np.median(ts):.3f} ms, p90 {np.percentile(ts, 90):.3f} ms, min {ts.min():.3f} ms", flush=True)
run("mono extract 1241x376 (obs_extract, pageable)", mono, {"pdl": 1, "graphs": 1})
run("mono extract, pdl off", mono, {"pdl": 0})
run("mono obs_extract, reused pageable arrays, plain enqueue, no pdl", mono_c, {"pdl": 0, "graphs": 0})
run("mono obs_extract, reused pageable arrays, plain enqueue, pdl", mono_c, {"pdl": 1, "graphs": 0})
run("mono obs_extract, reused pageable arrays, graph (default)", mono_c, {"pdl": 1, "graphs": 1})
run("stereo: two threads + obs_stereo_match (round-1 path)", stereo_threads, {"pdl": 1})
run("stereo: obs_stereo_frames, plain enqueue, no pdl", stereo_one_call, {"pdl": 0, "graphs": 0})
run("stereo: obs_stereo_frames, plain enqueue, pdl", stereo_one_call, {"pdl": 1, "graphs": 0})
run("stereo: obs_stereo_frames, graph, no pdl", stereo_one_call, {"pdl": 0, "graphs": 1})
run("stereo: obs_stereo_frames, graph + pdl (default)", stereo_one_call, {"pdl": 1, "graphs": 1})
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
