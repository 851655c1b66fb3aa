#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "contrastive" 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -25 > gpurun_out/pytest_tc.log; cat gpurun_out/pytest_tc.log | tail -25
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
