#!/bin/bash
# development (GPU box): time scratch variants with the split path on
cd "$(dirname "$0")/.."
for so in dryv_b200/csrc/libdryv_recon.so dryv_b200/csrc/libdryv_recon_var*.so; do
  [ -f $so ] || continue
  echo "== $so: $(cat ${so%.so}.flags 2>/dev/null)"
  for fr in ${FRAMES:-64 1}; do
    DRYV_SPLIT=1 DRYV_RECON_LIB=$PWD/$so timeout 300 python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'single', round(d['single_stream']['ms_per_step'],4), 'isolated', round(d['roofline']['kernel_ms_isolated'],4), 'parity', d['parity_vs_oracle_first_picture'])"
  done
done
