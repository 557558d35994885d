#!/bin/bash
# compute-sanitizer over the GPU suite (run on the GPU box through gpurun):
#     gpurun --timeout 2400 -- 'bash tools/sanitize.sh r02'
# memcheck over every GPU test; racecheck and synccheck over the extractor / stereo / BoW / object-layer tests (the searches with
# device-wide fixed-point iterations are too slow under racecheck).  Summaries land in gpurun_out/ for profiles/.
TAG=${1:-r02}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -m gpu -p no:cacheprovider"
$CS --tool memcheck --error-exitcode 1 --print-limit 20 $PY tests > gpurun_out/memcheck_${TAG}.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck_${TAG}.log
$CS --tool racecheck --error-exitcode 1 --print-limit 20 $PY tests/test_gpu_extractor.py tests/test_gpu_stereo.py tests/test_bow_matchers.py tests/test_frontend.py > gpurun_out/racecheck_${TAG}.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/racecheck_${TAG}.log
for f in memcheck racecheck; do
  { grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/${f}_${TAG}.log | tail -5; } > gpurun_out/${f}_${TAG}.txt
done
cat gpurun_out/memcheck_${TAG}.txt gpurun_out/racecheck_${TAG}.txt
