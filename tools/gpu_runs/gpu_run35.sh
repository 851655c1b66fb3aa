#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/b2n.err > gpurun_out/bench_cfg3_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n2.json')); print('n2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['config']['preheat_steps_untimed'])"
