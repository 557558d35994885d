set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_r01m.log 2>&1; tail -3 gpurun_out/pytest_r01m.log
timeout 300 python bench.py --workload extract > gpurun_out/bench_extract_r01m.json 2> gpurun_out/bench_extract_r01m.err; tail -c 600 gpurun_out/bench_extract_r01m.err
timeout 200 python bench.py --workload extract --impl reference --steps 3 --warmup 1 > gpurun_out/bench_extract_ref_r01m.json 2>&1
timeout 300 python bench.py > gpurun_out/bench_r01m.json 2> gpurun_out/bench_r01m.err
cut -c1-300 gpurun_out/bench_extract_r01m.json gpurun_out/bench_r01m.json
