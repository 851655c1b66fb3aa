#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n$n.json 2> gpurun_out/bench_cfg3_n$n.err; tail -2 gpurun_out/bench_cfg3_n$n.err | grep -v Warn | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n$n.json')); print('cfg3 N=$n', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
