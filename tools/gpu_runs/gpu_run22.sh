#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b1.err > gpurun_out/bench_cfg3_n1.json; tail -3 gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_n1.json')); print('cfg3 N=1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'])"
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b2.err > gpurun_out/bench_cfg2_n1.json; tail -3 gpurun_out/b2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_n1.json')); print('cfg2 N=1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'])"
