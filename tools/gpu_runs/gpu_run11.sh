#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python tools/parity_at_scale.py 2>&1 | grep -v Warn | tail -8
