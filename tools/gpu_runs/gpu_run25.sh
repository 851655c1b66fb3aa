#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -8 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | head -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1.json; tail -2 gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n1.json')); r=d['roofline']; print('cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', r['kernel_ms'], 'R', r['R'], 'R_emitted', r['R_emitted'])"
for w in 8 2; do
ISR_BWD_WARPS=$w timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b2.err > gpurun_out/bench_cfg2_w$w.json; tail -2 gpurun_out/b2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_w$w.json')); r=d['roofline']; print('cfg2 bwd warps=$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', r['kernel_ms'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_b.log 2>&1
