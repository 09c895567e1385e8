#!/bin/bash
# parity + A/B of the two decoder forms on small-vertex batches (one GPU)
TAG=${TAG:-r1g}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/check_rounds.py 2>&1 | tail -18
MOB200_ROUNDS=0 timeout 300 python tools/bench_small_vs.py > gpurun_out/${TAG}_small_rounds0.json 2> gpurun_out/${TAG}_small.err
timeout 300 python tools/bench_small_vs.py > gpurun_out/${TAG}_small_rounds_auto.json 2>> gpurun_out/${TAG}_small.err
cat gpurun_out/${TAG}_small_rounds0.json gpurun_out/${TAG}_small_rounds_auto.json | cut -c1-30,100-260
timeout 900 python tools/bench_configs.py > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
grep -c '"parity_ok": true' gpurun_out/${TAG}_configs.json; grep -c '"parity_ok": false' gpurun_out/${TAG}_configs.json
