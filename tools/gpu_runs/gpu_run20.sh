#!/usr/bin/env bash
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|MHz" | head -3
for gf in 1 0 1 0; do
ISR_GEOMETRY_FIRST=$gf python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err | tail -1 > gpurun_out/bench_cfg3_gf$gf.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_gf$gf.json')); print('cfg3 geomfirst=$gf', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
