#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in 8 2; do
ISR_FWD_WARPS=$w timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_w$w.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_w$w.json')); print('cfg3 warps=$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', d['roofline']['kernel_ms'])"
done
ISR_FWD_WARPS=2 timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b2.err > gpurun_out/bench_cfg2_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_n1.json')); print('cfg2 w2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', d['roofline']['kernel_ms'])"
