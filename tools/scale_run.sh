#!/bin/bash
# development: the bench line at N = 2, 4, 8 GPUs of one box (frame-sharded, no collective), launched the way the driver does
cd "$(dirname "$0")/.."
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err
done
