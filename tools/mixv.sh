#!/bin/bash
# development: tools/mix_probe.py for every scratch variant library
cd "$(dirname "$0")/.."
for so in dryv_b200/csrc/libdryv_recon_var*.so; do echo $so; DRYV_RECON_LIB=$so python tools/mix_probe.py ${1:-64} | tail -2; done
