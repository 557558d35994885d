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
run("stereo: two threads + obs_stereo_match (round-1 path)", stereo_threads, {"pdl": 1})
run("stereo: obs_stereo_frames, plain enqueue, no pdl", stereo_one_call, {"pdl": 0, "graphs": 0})
run("stereo: obs_stereo_frames, plain enqueue, pdl", stereo_one_call, {"pdl": 1, "graphs": 0})
run("stereo: obs_stereo_frames, graph, no pdl", stereo_one_call, {"pdl": 0, "graphs": 1})
run("stereo: obs_stereo_frames, graph + pdl (default)", stereo_one_call, {"pdl": 1, "graphs": 1})
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
