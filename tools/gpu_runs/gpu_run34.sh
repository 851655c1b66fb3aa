#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b4_$2.err > gpurun_out/bench_cfg3_n4_$2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n4_$2.json')); print('$2', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])"; }
run 29521 sampler
ISR_BENCH_NO_CLOCKS=1 run 29522 noclocks
run 29523 sampler2
