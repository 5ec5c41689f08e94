#!/bin/bash
# development (GPU box): time every scratch variant built by tools/variants.sh
cd "$(dirname "$0")/.."
for so in dryv_b200/csrc/libdryv_recon_var*.so; do
  echo "== $so: $(cat ${so%.so}.flags)"
  for fr in ${FRAMES:-64 16}; do
    DRYV_RECON_LIB=$PWD/$so timeout 300 python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'single', round(d['single_stream']['ms_per_step'],4), 'parity', d['parity_vs_oracle_first_picture'])"
  done
done
