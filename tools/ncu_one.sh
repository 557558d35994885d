#!/bin/bash
# One full ncu capture (source counters included) of one kernel of the stereo bench step, on the GPU box:
#     gpurun --timeout 900 -- 'bash tools/ncu_one.sh <kernel-regex> <tag> [launch-skip]'
# leaves gpurun_out/prof_<tag>.ncu-rep; read it with tools/ncu_summary.py / tools/ncu_segments.py
K=${1:-k_fast_cells}; TAG=${2:-x}; SKIP=${3:-8}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-sub --no-stress --frames 64 > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
