#!/bin/bash
# compute-sanitizer over the GPU suite (run on the GPU box through gpurun):
#     gpurun --timeout 2400 -- 'bash tools/sanitize.sh r02 [quick]'
# memcheck over every GPU test ("quick": only the extractor / stereo / C-ABI / drop-in tests); racecheck and synccheck over the
# extractor / stereo tests (the searches with device-wide fixed-point iterations are too slow under racecheck).
# Summaries land in gpurun_out/ for profiles/.
TAG=${1:-r02}
MODE=$2
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -m gpu -p no:cacheprovider"
CORE="tests/test_gpu_extractor.py tests/test_gpu_stereo.py tests/test_capi.py tests/test_dropin_cpp.py"
if [ "$MODE" = quick ]; then MEMSET="$CORE"; else MEMSET="tests"; fi
$CS --tool memcheck --error-exitcode 1 --print-limit 20 $PY $MEMSET > gpurun_out/memcheck_${TAG}.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck_${TAG}.log
# racecheck keeps host-side shadow state per allocation and kernel: one long pytest process grew past 200 GB and was killed by the
# kernel's OOM killer, so the tests run in several short processes
: > gpurun_out/racecheck_${TAG}.log
RC=0
for sel in "tests/test_gpu_extractor.py -k 'not many_seeds and not other_parameters'" "tests/test_gpu_extractor.py -k 'many_seeds or other_parameters'" \
           "tests/test_gpu_stereo.py -k 'not stereo_frames'" "tests/test_gpu_stereo.py -k 'stereo_frames_one_call'" "tests/test_gpu_stereo.py -k 'graph_replay'"; do
  eval "timeout 600 $CS --tool racecheck --error-exitcode 1 --print-limit 20 $PY $sel" >> gpurun_out/racecheck_${TAG}.log 2>&1 || RC=1
done
echo "racecheck rc=$RC" >> gpurun_out/racecheck_${TAG}.log
$CS --tool synccheck --error-exitcode 1 --print-limit 20 $PY tests/test_gpu_extractor.py tests/test_gpu_stereo.py > gpurun_out/synccheck_${TAG}.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/synccheck_${TAG}.log
for f in memcheck racecheck synccheck; do
  { echo "# tools/sanitize.sh $TAG $MODE: $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/${f}_${TAG}.log | tail -12; } > gpurun_out/${f}_${TAG}.txt
done
cat gpurun_out/memcheck_${TAG}.txt gpurun_out/racecheck_${TAG}.txt gpurun_out/synccheck_${TAG}.txt
