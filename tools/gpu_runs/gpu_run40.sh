#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric_fwd -s 5 -c 1 -o gpurun_out/prof_photo_fwd python tools/bench_next_rows.py > gpurun_out/ncu_photo.log 2>&1; tail -2 gpurun_out/ncu_photo.log
