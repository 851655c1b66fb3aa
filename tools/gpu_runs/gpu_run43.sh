#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -12 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -10
