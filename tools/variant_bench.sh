#!/bin/bash
# development: build kernel variants with different -D flags into scratch .so files and time them (bench line, 20 steps)
# usage: tools/variant_bench.sh "<flags A>" "<flags B>" ...
cd "$(dirname "$0")/.."
i=0
for flags in "$@"; do
  so=dryv_b200/csrc/libdryv_recon_var$i.so
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags \
      dryv_b200/csrc/recon.cu dryv_b200/csrc/recon_tables.cpp dryv_b200/csrc/levels_pack.cpp dryv_b200/csrc/cabac_host.cpp dryv_b200/csrc/multi.cpp -o $so 2>&1 | grep -E "error|spill"
  echo "== variant $i: $flags"
  for fr in 64 16; do
    DRYV_RECON_LIB=$so python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity_vs_oracle_first_picture'])"
  done
  i=$((i+1))
done
