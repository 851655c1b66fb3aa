#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err | grep -v Warn; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2.json')); print('cfg2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend ms', d['roofline']['kernel_ms'], 'ref', d.get('ref_cuda'))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --workload cfg2 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench2.log 2>&1
timeout 1200 python tools/parity_at_scale.py 2>&1 | grep -v Warn | tail -3 | cut -c1-200
