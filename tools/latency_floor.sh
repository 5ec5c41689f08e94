#!/bin/bash
# development: the wavefront kernel's time against the number of pictures in the launch — one picture is the dependency
# chain alone (W + 2H hops), many pictures the throughput. Run on a GPU box.
cd "$(dirname "$0")/.."
for fr in 1 2 4 8 16 32 64 128; do
  timeout 300 python bench.py --steps 20 --warmup 3 --frames $fr --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames', $fr, 'in flight ms/step', round(d['ms_per_step'],4), 'single stream ms', round(d['single_stream']['ms_per_step'],4))"
done
