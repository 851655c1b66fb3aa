#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_cfg3_n2.err; tail -3 gpurun_out/bench_cfg3_n2.err; cat gpurun_out/bench_cfg3_n2.json
