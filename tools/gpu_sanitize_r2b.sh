#!/bin/bash
# compute-sanitizer over the run-major / chained-round paths of block mode.  TAG=r2x bash tools/gpu_sanitize_r2b.sh
TAG=${TAG:-r2}
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_blockmode.py tests/test_gpu_parity.py -m gpu -q -x -k "rounds" 2>&1 | tail -8 > gpurun_out/${TAG}_memcheck_rounds.log; echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/${TAG}_memcheck_rounds.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_blockmode.py -m gpu -q -x -k "rounds_run_major and (quat12 or exp15 or color8) or rounds_mixed or unit_chains or stale or error_streams" 2>&1 | tail -8 > gpurun_out/${TAG}_racecheck_rounds.log; echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/${TAG}_racecheck_rounds.log
cat gpurun_out/${TAG}_memcheck_rounds.log gpurun_out/${TAG}_racecheck_rounds.log
