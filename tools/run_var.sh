#!/bin/bash
# development: time scratch variants dryv_b200/csrc/libdryv_recon_var*.so (and the main library) with the bench line
cd "$(dirname "$0")/.."
for so in dryv_b200/csrc/libdryv_recon.so dryv_b200/csrc/libdryv_recon_var*.so; do
  for fr in 64 16; do
    DRYV_RECON_LIB=$so python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$so frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity_vs_oracle_first_picture'])"
  done
done
