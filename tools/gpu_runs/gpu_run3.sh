#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err; cat gpurun_out/bench_cfg3.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
