#!/bin/bash
# usage: tools/profile_k.sh <tag> <kernel-regex> [count]   (run on the GPU box)
TAG=$1; RE=$2; CNT=${3:-3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 24 -c $CNT -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
