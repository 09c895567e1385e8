#!/bin/bash
# fused / walk-only / decode-only timing of the bench workload (diagnostics); DEBUG=1 adds a pass with cycle counters
V=${VERTS:-67108864}; S=${SEG:-4096}
if [ "$DEBUG" = "1" ]; then
  MOB200_DEBUG_COUNTERS=1 python -c "from meshoptimizer_b200 import build; build.build(force=True)"
  python tools/gpu_diag.py $V $S 10
  MOB200_WALKER_LEAD=4294967295 python tools/gpu_diag.py $V $S 10
  MOB200_WALKER_LEAD=4294967294 python tools/gpu_diag.py $V $S 10
  python -c "from meshoptimizer_b200 import build; build.build(force=True)"
fi
python tools/gpu_diag.py $V $S 10 | tail -1
MOB200_WALKER_LEAD=4294967295 python tools/gpu_diag.py $V $S 10 | tail -1
MOB200_WALKER_LEAD=4294967294 python tools/gpu_diag.py $V $S 10 | tail -1
