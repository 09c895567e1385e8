#!/bin/bash
# quick perf check: decode-bound (256-vertex streams) and mixed (4096-vertex streams) at 16 Mi vertices
for seg in 256 4096; do
  timeout 300 python bench.py --verts 16777216 --steps 10 --warmup 3 --segment $seg --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['streams_per_gpu'], 'streams: value', round(d['value']), 'GB/s decoded, roofline frac', round(d['roofline']['frac'],3), 'ms', round(d['ms_per_step'],3))"
done
