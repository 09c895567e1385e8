#!/bin/bash
# A/B of variant builds: fused time (CUDA events) and DRAM bytes of one launch (ncu) on the bench workload
for so in meshoptimizer_b200/lib/variants/*.so; do
  echo "== $(basename $so .so)"
  MOB200_LIB=$PWD/$so python tools/gpu_diag.py 67108864 4096 10 | tail -1 | sed 's/verts.*env//'
  MOB200_LIB=$PWD/$so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:decode_kernel -s 3 -c 1 python tools/gpu_diag.py 67108864 4096 2 2>&1 | grep -E "dram__bytes|gpu__time"
done
