#!/bin/bash
# Run on the GPU box (via gpurun): launch list of one bench run + full captures of each kernel.
# usage: tools/profile.sh <tag>
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 64"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
# one full capture per kernel (skip the warm-up launches: 3 warmup steps x 24 launches)
ncu --set full --clock-control none --import-source on -k regex:'k_fast_cells|k_quadtree|k_blur|k_describe|k_stereo_match|k_stereo_filter' -s 10 -c 10 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/prof_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_resize' -s 21 -c 7 -o gpurun_out/prof_resize_${TAG} -f $CMD > gpurun_out/prof_resize_${TAG}.log 2>&1
ls -la gpurun_out
