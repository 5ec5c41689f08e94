#!/bin/bash
# development: tools/overlap_probe.py (1..4 batches in flight) for the main library and every scratch variant
cd "$(dirname "$0")/.."
for so in dryv_b200/csrc/libdryv_recon.so dryv_b200/csrc/libdryv_recon_var*.so; do
  for fr in 64 16; do
    echo -n "$so: "; DRYV_RECON_LIB=$so python tools/overlap_probe.py $fr 20 2>&1 | tail -1
  done
done
