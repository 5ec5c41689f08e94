for v in "" _var0 _var1 _var2; do
  for fr in 64 16; do
    DRYV_RECON_LIB=dryv_b200/csrc/libdryv_recon$v.so python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib$v frames', $fr, 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity_vs_oracle_first_picture'])"
  done
done
