#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out/golden
timeout 300 python tests/golden/make_tracker_golden.py gpurun_out/golden 2>&1 | tail -3
cp gpurun_out/golden/tracker_g1.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -15 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | head -10
timeout 600 python tools/bench_next_rows.py 2> gpurun_out/next.err > gpurun_out/next_rows.json; tail -3 gpurun_out/next.err; cat gpurun_out/next_rows.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_blend_fwd_v4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
