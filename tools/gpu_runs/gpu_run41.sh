#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -15 > gpurun_out/pytest_gpu.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu.log | head -10
ISR_PHOTO_TILE=32 timeout 300 python -m pytest tests/test_losses.py -m gpu -x -q 2>&1 | grep -E "passed|failed|rror" | head -3
for t in 16 32; do ISR_PHOTO_TILE=$t timeout 300 python tools/bench_next_rows.py 2>/dev/null > gpurun_out/next_rows_$t.json; python -c "
import json,sys; d=json.load(open('gpurun_out/next_rows_$t.json'))['photometric_1080p']; print('tile $t', {k:round(v,4) for k,v in d.items() if 'ms' in k or 'GBs' in k})"; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
