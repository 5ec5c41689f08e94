#!/bin/bash
# development: build kernel variants with different -D flags into scratch .so files HERE (no GPU needed); tools/variants_run.sh times them on a GPU box
# usage: tools/variants.sh "<flags A>" "<flags B>" ...
cd "$(dirname "$0")/.."
rm -f dryv_b200/csrc/libdryv_recon_var*.so
i=0
for flags in "$@"; do
  so=dryv_b200/csrc/libdryv_recon_var$i.so
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags -Xptxas -v \
      dryv_b200/csrc/recon.cu dryv_b200/csrc/recon_tables.cpp dryv_b200/csrc/levels_pack.cpp dryv_b200/csrc/cabac_host.cpp dryv_b200/csrc/multi.cpp -o $so 2>&1 | grep -A3 "recon_wavefront" | grep -E "error|spill|registers"
  echo "$flags" > dryv_b200/csrc/libdryv_recon_var$i.flags
  python tools/sass_hist.py $so wavefront | head -1
  i=$((i+1))
done
