#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -2 gpurun_out/bench_cfg3.err | grep -v Warn; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3.json')); print('cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend ms', d['roofline']['kernel_ms'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 220 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1
