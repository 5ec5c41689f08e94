#!/bin/bash
# development: parity subset + bench numbers of the default library (or DRYV_RECON_LIB); run on a GPU box
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for fr in 64 16; do
  timeout 300 python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'single', round(d['single_stream']['ms_per_step'],4), 'isolated', round(d['roofline']['kernel_ms_isolated'],4), 'parity', d['parity_vs_oracle_first_picture'])"
done
