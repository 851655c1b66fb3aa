#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/b2p.err > gpurun_out/bench_cfg3_n2_prefetch.json; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n2_prefetch.json')); print('n2 prefetch value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['geometry_prefetch'][:20])"
