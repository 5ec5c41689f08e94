#!/bin/bash
# development (GPU box): row teams per launch (DRYV_WAVE_GRID) against the default sm_count * 8
cd "$(dirname "$0")/.."
python tools/mix_probe.py 64
for g in 0 1088 1036 888 1184; do
  DRYV_WAVE_GRID=$g timeout 300 python bench.py --steps 20 --warmup 3 --frames 64 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('grid', $g, 'ms/step', round(d['ms_per_step'],4), 'single', round(d['single_stream']['ms_per_step'],4), 'isolated', round(d['roofline']['kernel_ms_isolated'],4), 'parity', d['parity_vs_oracle_first_picture'])"
done
