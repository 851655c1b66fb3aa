#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out/golden
timeout 200 python tests/golden/make_contrastive_golden.py gpurun_out/golden 2>&1 | tail -2
cp gpurun_out/golden/contrastive_g1.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -15 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1.json ) 2>&1 | grep real; tail -2 gpurun_out/b1.err; cat gpurun_out/bench_cfg3_n1.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/b3.err > gpurun_out/bench_ref.json ) 2>&1 | grep real; cat gpurun_out/bench_ref.json | cut -c1-300
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tracker.py tests/test_losses.py tests/test_contrastive_golden.py -m gpu -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -10
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_b2.log 2>&1
