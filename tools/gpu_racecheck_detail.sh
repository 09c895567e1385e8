#!/bin/bash
# racecheck with the hazard list: one rounds case and one multi-block case of the plain form.
# MOB200_LIB=.../variants/arriveall.so (tools/ab_build.py arriveall=MOB200_X_ARRIVE_ALL) repeats it with an arrival per lane on tile_free.
TAG=${TAG:-r1k}
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "rounds_small and 1000000" 2>&1 | grep -v "^=========     Saved\|^$" | cut -c1-330 | head -120 > gpurun_out/${TAG}_racecheck_rounds_detail.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "c2_segments and 4096" 2>&1 | grep -v "^$" | cut -c1-330 | head -120 > gpurun_out/${TAG}_racecheck_plain_detail.log
grep -h "RACECHECK SUMMARY\|passed\|failed" gpurun_out/${TAG}_racecheck_rounds_detail.log gpurun_out/${TAG}_racecheck_plain_detail.log
