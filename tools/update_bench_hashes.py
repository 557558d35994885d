#!/usr/bin/env python
"""Writes tests/golden/bench_hashes.json from the JSON line of a ONE-GPU bench.py run: the result hashes of the sharded batch
workloads (configs[3], configs[4]) that every multi-GPU run must reproduce.

    python bench.py > gpurun_out/bench.json ; python tools/update_bench_hashes.py gpurun_out/bench.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
line = None
for l in open(sys.argv[1]):
    l = l.strip()
    if l.startswith("{") and '"metric"' in l:
        line = json.loads(l)
assert line is not None and line.get("n_gpus") == 1, "need the JSON line of a 1-GPU run"
out_path = os.path.join(ROOT, "tests", "golden", "bench_hashes.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
subs = line.get("sub") or {"_": line}
for s in subs.values():
    if not s or "result_hash" not in s:
        continue
    wl = s["config"]["workload"]
    if wl.startswith("configs[3]"):
        T = int(wl.split("extraction of ")[1].split()[0])
        pool = int(s["config"]["pool"].split()[0])
        out[f"extract:T={T}:pool={pool}"] = s["result_hash"]["value"]
    elif wl.startswith("configs[4]"):
        K = int(wl.split("configs[4]: ")[1].split()[0])
        Wn = int(wl.split("+-")[1].split()[0])
        out[f"knn2:K={K}:W={Wn}"] = s["result_hash"]["value"]
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out))
