#!/bin/bash
# A/B timing of variant builds (meshoptimizer_b200/lib/variants/*.so, made by tools/ab_build.py) on the GPU box:
#   fused, walk-only and decode-only time of the bench workload per variant
V=${VERTS:-67108864}; S=${SEG:-4096}
for so in meshoptimizer_b200/lib/variants/*.so; do
  echo "== $(basename $so .so)"
  MOB200_LIB=$PWD/$so python tools/gpu_diag.py $V $S 10 | tail -1 | sed 's/verts.*env//'
  MOB200_LIB=$PWD/$so MOB200_WALKER_LEAD=4294967295 python tools/gpu_diag.py $V $S 10 | tail -1 | sed 's/verts.*env//'
  MOB200_LIB=$PWD/$so MOB200_WALKER_LEAD=4294967294 python tools/gpu_diag.py $V $S 10 | tail -1 | sed 's/verts.*env//'
done
