#!/bin/bash
# Run on the GPU box (via gpurun): launch list of one bench run + one full ncu capture (with source counters) of every kernel of a step.
# usage: tools/profile.sh <tag>          -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep
# then here:  tools/ncu_summary.py launches / full ...  > profiles/...;  tools/ncu_pipes.py 64 profiles/traffic.json gpurun_out/prof_<tag>.ncu-rep
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub --no-stress --frames 64"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
# a step is 2 x 8 extraction launches + 3 stereo launches per pipeline; 40 consecutive launches after the warm-up hold every kernel and grid
# (gpurun copies back at most 64 MiB: 36 launches with source counters stay below it)
ncu --set full --clock-control none --import-source on -k regex:'k_' -s 120 -c 36 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/launches_${TAG}.csv gpurun_out/prof_${TAG}.ncu-rep
