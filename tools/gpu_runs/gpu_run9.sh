#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -2 gpurun_out/bench_cfg3.err | grep -v Warn; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3.json')); print('cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend ms', d['roofline']['kernel_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_blend_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extra_sparse_bwd -s 3 -c 1 -o gpurun_out/prof_sparse_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_s.log 2>&1
