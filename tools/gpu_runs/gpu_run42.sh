#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_semantic_step.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -12 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | head -10
for pf in 1 0; do ISR_BENCH_PREFETCH=$pf timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b_pf$pf.err > gpurun_out/bench_cfg3_pf$pf.json; tail -1 gpurun_out/b_pf$pf.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_pf$pf.json')); print('prefetch=$pf cfg3', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"; done
