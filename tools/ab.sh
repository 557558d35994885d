#!/bin/bash
# A/B of library variants on one box: tools/ab.sh <rounds> <variant> [<variant> ...]   (variants/lib_<variant>.so, built beforehand
# with tools/mkvariant.sh).  Runs the headline bench with every variant in turn, <rounds> times; one line per run is appended to
# gpurun_out/ab.log (value / e2e / stress / per-kernel serialised times in us per 128 images).
R=$1; shift
mkdir -p gpurun_out
cp object_slam_b200/libobslam_b200.so /tmp/lib_orig.so
for r in $(seq $R); do
  for v in "$@"; do
    cp variants/lib_$v.so object_slam_b200/libobslam_b200.so
    python bench.py --no-cpu-baseline --no-sub ${AB_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
p=d['roofline']['per_launch_ms']
print('$v', 'value %.0f e2e %.0f stress %.0f' % (d['value'], d['e2e']['value'], d.get('stress', {}).get('value', 0)), ' '.join('%s %.0f' % (k, 1e3*v) for k,v in p.items()), flush=True)"
  done
done | tee -a gpurun_out/ab.log
cp /tmp/lib_orig.so object_slam_b200/libobslam_b200.so
