"""Times obs_hamming_knn2 with the POPC engine and both tensor-core variants on one GPU (device-resident inputs, CUDA events) and checks that their
outputs are identical.  Usage: python tools/knn2_probe.py [keyframes] [window] [n_desc]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_slam_b200 import sharding  # noqa: E402
from object_slam_b200._capi import check, lib  # noqa: E402
from object_slam_b200.matcher import ORBmatcher  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Wn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)
D = torch.randint(0, 256, (K, N, 32), dtype=torch.uint8, device=dev, generator=g)
m = int(N * 0.3)
for k in range(1, K):
    noise = torch.randint(0, 256, (3, m, 32), dtype=torch.uint8, device=dev, generator=g)
    D[k, N - m:] = D[k - 1, :m] ^ (noise[0] & noise[1] & noise[2])
pairs = torch.from_numpy(sharding.window_pairs(0, K, K, Wn)).to(dev)
P = pairs.shape[0]
M = ORBmatcher(0.6, True, device=0)
out = {}
res = {}
for name, eng, pair in (("popc", M.KNN2_POPC, 1), ("tensor", M.KNN2_TENSOR, 0), ("tensor_cta_pair", M.KNN2_TENSOR, 1)):
    M.set_knn2_engine(eng)
    check(lib().obs_set_option(b"knn2_cta_pair", pair))
    bi = torch.empty((P, N), dtype=torch.int32, device=dev)
    bd = torch.empty((P, N), dtype=torch.int32, device=dev)
    sd = torch.empty((P, N), dtype=torch.int32, device=dev)
    def run():
        check(lib().obs_hamming_knn2(M._h, C.c_void_p(D.data_ptr()), K, N, C.c_void_p(pairs.data_ptr()), P, 50, C.c_float(0.6),
                                     C.c_void_p(bi.data_ptr()), C.c_void_p(bd.data_ptr()), C.c_void_p(sd.data_ptr())))
    st = torch.cuda.ExternalStream(M.stream)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(st)
    for _ in range(reps):
        run()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name] = {"ms": ms, "T_dist_per_s": P * N * N / (ms * 1e-3) / 1e12}
    res[name] = (bi.cpu().numpy(), bd.cpu().numpy(), sd.cpu().numpy())
same = all(np.array_equal(a, b) and np.array_equal(a, c) for a, b, c in zip(res["popc"], res["tensor"], res["tensor_cta_pair"]))
out["identical"] = bool(same)
out["config"] = {"keyframes": K, "window": Wn, "n_desc": N, "pairs": int(P), "matches": int((res["popc"][0] >= 0).sum())}
if not same:
    for a, b, nm in zip(res["popc"], res["tensor"], ("best_idx", "best_dist", "second_dist")):
        bad = np.argwhere(a != b)
        out["mismatch_" + nm] = {"count": int(len(bad)), "first": [[int(x) for x in r] + [int(a[tuple(r)]), int(b[tuple(r)])] for r in bad[:8]]}
print(json.dumps(out))
