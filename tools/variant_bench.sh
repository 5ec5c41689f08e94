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
  DRYV_RECON_LIB=$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms', d['value'], 'Mpx/s', d['parity_vs_oracle_first_picture'])"
  DRYV_RECON_LIB=$so python bench.py --steps 10 --warmup 3 --frames 16 --no-cpu-baseline --no-e2e --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  16 frames:', d['ms_per_step'], 'ms', d['parity_vs_oracle_first_picture'])"
  i=$((i+1))
done
