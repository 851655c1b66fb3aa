#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out/golden
timeout 300 python tests/golden/make_tracker_golden.py gpurun_out/golden 2>&1 | tail -3
cp gpurun_out/golden/tracker_g1.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -15 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | head -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1.json; tail -2 gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n1.json')); r=d['roofline']; print('cfg3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'blend_ms', r['kernel_ms'])"
( time timeout 900 python bench.py --workload cfg5 --steps 10 --warmup 3 2> gpurun_out/b5.err > gpurun_out/bench_cfg5_n1.json ) 2>&1 | grep real; tail -3 gpurun_out/b5.err; cat gpurun_out/bench_cfg5_n1.json
