#!/bin/bash
# parity + A/B of the two decoder forms on small-vertex batches (one GPU)
TAG=${TAG:-r1f}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json; echo
MOB200_ROUNDS=0 timeout 300 python tools/bench_small_vs.py > gpurun_out/${TAG}_small_rounds0.json 2> gpurun_out/${TAG}_small.err
MOB200_ROUNDS=1 timeout 300 python tools/bench_small_vs.py > gpurun_out/${TAG}_small_rounds1.json 2>> gpurun_out/${TAG}_small.err
cat gpurun_out/${TAG}_small_rounds0.json gpurun_out/${TAG}_small_rounds1.json | cut -c1-200
