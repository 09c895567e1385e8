#!/bin/bash
# round-end measurement pass (one GPU): parity tests, bench lines, ncu launch list + full capture, config sweep
TAG=${TAG:-r1d}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_full python tools/gpu_diag.py 67108864 4096 2 > gpurun_out/${TAG}_ncu_c.log 2>&1
timeout 900 python tools/bench_configs.py > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" > gpurun_out/${TAG}_lscpu.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
