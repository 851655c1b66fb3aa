#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 900 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err | grep -v Warn; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2.json')); print('cfg2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
