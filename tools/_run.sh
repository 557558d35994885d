( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
( time timeout 300 python bench.py ) > gpurun_out/bench_r01p.json 2> gpurun_out/bench_r01p.err; tail -4 gpurun_out/bench_r01p.err; cut -c1-260 gpurun_out/bench_r01p.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
