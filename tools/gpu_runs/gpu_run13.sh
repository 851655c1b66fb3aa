#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_blend_bwd python bench.py --steps 2 --warmup 3 --workload cfg2 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
