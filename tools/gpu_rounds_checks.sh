#!/bin/bash
# rounds form: parity tests, many-stream small-vertex throughput with every byte compared, the forced form on few long streams
TAG=${TAG:-r1s}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/bench_small_vs.py > gpurun_out/${TAG}_small.json 2>/dev/null; python -c "
import sys,json
for l in open('gpurun_out/${TAG}_small.json'):
    d=json.loads(l); print('  vs',d['vertex_size'],d['streams'],'%.4f ms %.0f GB/s'%(d['best_ms'],d['decoded_GBps']),d['ok'])"
timeout 300 python tools/check_rounds.py 2>&1 | tail -2
