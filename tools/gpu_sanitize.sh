#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (memcheck: all kernels; racecheck: shared-memory hazards)
SEL='kat or fixture or index or meshlet or gltf or error or tail or small or unaligned or guard'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" 2>&1 | tail -6 > gpurun_out/r1d_memcheck.log; echo "memcheck rc=$?" >> gpurun_out/r1d_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "kat or index or meshlet or gltf" 2>&1 | tail -6 > gpurun_out/r1d_racecheck.log; echo "racecheck rc=$?" >> gpurun_out/r1d_racecheck.log
cat gpurun_out/r1d_memcheck.log gpurun_out/r1d_racecheck.log
