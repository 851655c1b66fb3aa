#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -15 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1.json; tail -2 gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n1.json')); r=d['roofline']; print('cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', r['kernel_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contrast_loss_tc -s 3 -c 1 -o gpurun_out/prof_contrast_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_tc.log 2>&1; tail -2 gpurun_out/ncu_full_tc.log
