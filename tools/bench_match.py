"""Device-resident timing of the matcher kernels on one GPU (run on the B200 box):
popc ceilings, brute-force keyframe matching (cfg5 shape), projection search (cfg3 shape)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
from object_slam_b200 import synth  # noqa: E402
from object_slam_b200._capi import lib, check  # noqa: E402
from object_slam_b200.matcher import ORBmatcher  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for mode, name in ((0, "plain"), (1, "csa5"), (3, "csa4"), (2, "popc_only")):
    g = C.c_double()
    check(lib().obs_microbench_popc(0, mode, C.byref(g)))
    out[f"popc_peak_{name}_gdist_s"] = g.value

M = ORBmatcher(0.6, True)
st = torch.cuda.ExternalStream(M.stream)


def timed(fn, reps=5):
    fn(); M.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    M.sync()
    return e0.elapsed_time(e1) / reps


# ---- knn2 (cfg5 shape: 2000 descriptors per keyframe)
K, n = int(os.environ.get("KNN_K", 96)), 2000
D = torch.from_numpy(synth.keyframe_descriptors(K, n, 0)).to(dev)
pairs = np.array([(i, j) for i in range(K) for j in range(K) if i != j], np.int32)
P = len(pairs)
dp = torch.from_numpy(pairs).to(dev)
bi = torch.empty((P, n), dtype=torch.int32, device=dev)
ms = timed(lambda: M.knn2(D.data_ptr(), dp.data_ptr(), n_keyframes=K, n_desc=n, best_idx=bi.data_ptr()) if False else
           check(lib().obs_hamming_knn2(M._h, C.c_void_p(D.data_ptr()), K, n, C.c_void_p(dp.data_ptr()), P, 50, 0.6,
                                        C.c_void_p(bi.data_ptr()), None, None)))
out["knn2"] = {"keyframes": K, "pairs": P, "ms": ms, "gdist_s": P * n * n / (ms * 1e-3) / 1e9,
               "queries_s": P * n / (ms * 1e-3), "matches": int((bi >= 0).sum())}

# ---- projection search (cfg3 shape)
B, NK, NM = int(os.environ.get("PROJ_B", 128)), 1000, 20000
import matcher_cases as mc  # noqa: E402
M.mfNNratio = 0.8
cases = [mc.map_case(synth.TUM_SHAPE, NK, NM, s) for s in range(8)]
frames = [cases[b % 8][0] for b in range(B)]
fs = M.frame_set(synth.scale_factors(), mc.bounds(synth.TUM_SHAPE), synth.camera_for(synth.TUM_SHAPE), max_frames=B, max_keypoints=NK)
t0 = time.perf_counter(); fs.upload(frames); M.sync(); out["frame_set_upload_ms_per_frame"] = (time.perf_counter() - t0) * 1e3 / B
arr = [torch.from_numpy(np.stack([cases[b % 8][1][k] for b in range(B)])).to(dev) for k in mc.MP_KEYS]
kpm = torch.empty((B, fs.cap), dtype=torch.int32, device=dev)
nm = torch.empty(B, dtype=torch.int32, device=dev)
ms = timed(lambda: M.SearchByProjection(fs, *[a.data_ptr() for a in arr], th=3.0, n_points=NM, per_frame=True,
                                        kp_match=kpm.data_ptr(), n_matches=nm.data_ptr()))
# host-output call once to read the round counts
n_h, match_h = M.SearchByProjection(fs, *[a.data_ptr() for a in arr], th=3.0, n_points=NM, per_frame=True)
out["projection"] = {"frames": B, "points": NM, "keypoints": NK, "ms": ms, "frames_s": B / (ms * 1e-3),
                     "points_s": B * NM / (ms * 1e-3), "rounds_mean": float(M.last_rounds().mean()),
                     "rounds_max": int(M.last_rounds().max()), "matches_mean": float(n_h.mean())}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_match.json"), "w"), indent=1)
