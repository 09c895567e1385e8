#!/bin/bash
# compute-sanitizer over batches in which a unit decodes several blocks (tile and slot reuse), both forms of the decoders:
# the rounds form forced on (rounds_small, rounds_mixed) and the plain form (c2_segments, c4): memcheck, then racecheck
TAG=${TAG:-r1n}
SEL="rounds_small or rounds_mixed or c2_segments or c4_small"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" 2>&1 | tail -6 > gpurun_out/${TAG}_memcheck_multiblock.log; echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck_multiblock.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" 2>&1 | grep -v "^$" | cut -c1-300 | tail -30 > gpurun_out/${TAG}_racecheck_multiblock.log; echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck_multiblock.log
cat gpurun_out/${TAG}_memcheck_multiblock.log; tail -8 gpurun_out/${TAG}_racecheck_multiblock.log
