#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -12 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | head -10
( time timeout 900 python bench.py 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1.json ) 2>&1 | grep real; tail -2 gpurun_out/b1.err; cat gpurun_out/bench_cfg3_n1.json
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/b2.err > gpurun_out/bench_cfg2_n1.json; tail -2 gpurun_out/b2.err; cat gpurun_out/bench_cfg2_n1.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/b3.err > gpurun_out/bench_ref.json ) 2>&1 | grep real; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_b.log 2>&1
