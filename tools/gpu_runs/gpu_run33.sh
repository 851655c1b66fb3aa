#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 2> gpurun_out/b4.err > gpurun_out/bench_cfg3_n4.json; tail -3 gpurun_out/b4.err; cat gpurun_out/bench_cfg3_n4.json | cut -c1-900
