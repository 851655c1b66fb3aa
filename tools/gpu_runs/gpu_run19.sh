#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err | tail -1 > gpurun_out/bench_cfg3_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n1.json')); print('cfg3 N=1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_cfg3_n2.err; tail -3 gpurun_out/bench_cfg3_n2.err | grep -v Warn | cut -c1-300; wc -l gpurun_out/bench_cfg3_n2.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_n2.json').read().strip().splitlines()[-1]); print('cfg3 N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
