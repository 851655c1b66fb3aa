#!/usr/bin/env bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_losses.py -m gpu -x -q 2>&1 | grep -E "passed|failed|rror" | head
ISR_PHOTO_TILE=32 timeout 300 python -m pytest tests/test_losses.py -m gpu -x -q 2>&1 | grep -E "passed|failed|rror" | head
for t in 16 32; do ISR_PHOTO_TILE=$t timeout 300 python tools/bench_next_rows.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)['photometric_1080p']; print('tile $t', {k:round(v,4) for k,v in d.items() if 'ms' in k or 'GBs' in k})"; done
