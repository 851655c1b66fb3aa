#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -15 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -10
timeout 600 python tools/bench_next_rows.py 2> gpurun_out/next.err > gpurun_out/next_rows.json; tail -3 gpurun_out/next.err; cat gpurun_out/next_rows.json
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/b2.err > gpurun_out/bench_cfg2_n1.json; tail -2 gpurun_out/b2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_n1.json')); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', r['kernel_ms'], d.get('ref_cuda'))"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1_short.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n1_short.json')); print('cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['views'])"
