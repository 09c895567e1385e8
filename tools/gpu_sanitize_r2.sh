#!/bin/bash
# compute-sanitizer over the round-2 paths: block mode (sidecars incl. stale / garbage ones), the eight-lane meshlet
# kernel, the error / garbage vertex streams, the glTF bound checks.  TAG=r2x bash tools/gpu_sanitize_r2.sh
TAG=${TAG:-r2}
SEL='stale or garbage or changed or error_streams or needs_offsets or segmenter or meshlet or gltf or error_codes or kat or reuses_offsets and (c4 or c2_mono_ragged or c3_quat12_v0)'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_blockmode.py tests/test_gpu_index_gltf.py tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" 2>&1 | tail -8 > gpurun_out/${TAG}_memcheck.log; echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/${TAG}_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_blockmode.py tests/test_gpu_index_gltf.py -m gpu -q -x -k "meshlet or stale or error_streams or (reuses_offsets and c2_mono_ragged)" 2>&1 | tail -8 > gpurun_out/${TAG}_racecheck.log; echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/${TAG}_racecheck.log
cat gpurun_out/${TAG}_memcheck.log gpurun_out/${TAG}_racecheck.log
