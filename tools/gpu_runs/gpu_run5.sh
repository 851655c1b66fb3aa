#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err; cut -c1-400 gpurun_out/bench_cfg3.json
timeout 900 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --workload cfg2 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_blend_bwd python bench.py --steps 2 --warmup 3 --workload cfg2 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
