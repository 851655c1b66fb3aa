#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/bench_next_rows.py 2> gpurun_out/next.err > gpurun_out/next_rows.json; tail -3 gpurun_out/next.err; cat gpurun_out/next_rows.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"photometric|tracker" -c 40 --csv --log-file gpurun_out/launches_next.csv python tools/bench_next_rows.py > gpurun_out/ncu_next.log 2>&1; grep -c photometric gpurun_out/launches_next.csv
