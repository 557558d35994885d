#!/bin/bash
# Quick A/B loop on the GPU box: extractor + stereo parity tests, the headline bench line without sub-results, optionally one ncu capture.
#     gpurun --timeout 900 -- 'bash tools/quick.sh <tag> [kernel-regex]'
TAG=${1:-x}; K=$2
mkdir -p gpurun_out
python -m pytest tests/test_gpu_extractor.py tests/test_gpu_stereo.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-sub > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.0f e2e %.0f stress %.0f" % (d["value"], d["e2e"]["value"], d["stress"]["value"]))
print({k: round(v, 4) for k, v in d["roofline"]["per_launch_ms"].items()})
PY
if [ -n "$K" ]; then bash tools/ncu_one.sh "$K" ${TAG}_ncu; fi
