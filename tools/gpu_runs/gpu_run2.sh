#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tests/golden/make_golden.py gpurun_out/golden 2>&1 | tail -8
timeout 600 python tools/quick_bench.py 2>&1 | tail -5
