#!/bin/bash
# development (GPU box): the split path (DRYV_SPLIT=1) — parity subset + bench timing beside the row-team kernel
cd "$(dirname "$0")/.."
DRYV_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
for sp in 1 0; do
for fr in ${FRAMES:-64 16 1}; do
  DRYV_SPLIT=$sp timeout 300 python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split', $sp, 'frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'single', round(d['single_stream']['ms_per_step'],4), 'isolated', round(d['roofline']['kernel_ms_isolated'],4), 'parity', d['parity_vs_oracle_first_picture'])"
done
done
