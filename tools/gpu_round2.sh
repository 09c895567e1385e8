#!/bin/bash
# Round-2 GPU pass on one B200: bench both arms, ncu launch list + one --set full capture of the headline kernel.
#   TAG=r2c bash tools/gpu_round2.sh
TAG=${TAG:-r2}
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" > $O/${TAG}_lscpu.txt 2>&1
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 1200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
if [ -n "$NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu > $O/${TAG}_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o $O/${TAG}_full python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-e2e > $O/${TAG}_ncu_c.log 2>&1
fi
tail -c 600 $O/${TAG}_bench.err
if [ -n "$EXTRA" ]; then
timeout 300 python tools/bench_meshlet.py > $O/${TAG}_meshlet.json 2> $O/${TAG}_meshlet.err
timeout 300 python tools/bench_filters.py > $O/${TAG}_filters.json 2> $O/${TAG}_filters.err
timeout 300 python tools/bench_index.py > $O/${TAG}_index.json 2> $O/${TAG}_index.err
true
fi
