#!/bin/bash
# development: build kernel variants with different -D flags into scratch .so files and time them
# usage: tools/variant_bench.sh "<flags A>" "<flags B>" ...
cd "$(dirname "$0")/.."
i=0
for flags in "$@"; do
  so=dryv_b200/csrc/libdryv_recon_var$i.so
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags \
      dryv_b200/csrc/recon.cu dryv_b200/csrc/recon_tables.cpp -o $so 2>&1 | grep -E "error|spill" 
  echo "== variant $i: $flags"
  for fr in 64 16; do
  DRYV_RECON_LIB=$so ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv python bench.py --steps 2 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep -o '"dryv::[a-z_]*.*' | awk -F'"' -v fr=$fr '{n[$2]++; s[$2]+=$(NF-1)} END {for (k in n) printf "  %d frames  %s  %.1f us\n", fr, k, s[k]/n[k]/1000}'
  done
  DRYV_RECON_LIB=$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  bench 64:', d['ms_per_step'], 'ms', d['value'], 'Mpx/s', d['parity_vs_oracle_first_picture'])"
  i=$((i+1))
done
