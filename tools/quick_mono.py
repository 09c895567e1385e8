"""first-time decode (serial path) of one monolithic C2 stream and of the C1a grid: python tools/quick_mono.py [verts]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import meshoptimizer_b200 as mb
from oracle import workloads
from tests.gpu_util import device_run
verts = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
for w in (workloads.c2(total=verts, seg=None), workloads.c2(total=verts, seg=1 << 16), workloads.c1a(version=1, level=2), workloads.c1a(version=0, level=0)):
    outs, status, plan, guard = device_run(w, runs=3)
    ms = min(plan.timing_history(3))
    ok = bool((status == 0).all() and guard and np.array_equal(np.concatenate(outs), w.source))
    print(json.dumps({"workload": w.name, "streams": w.n, "serial_ms": ms, "GBps": w.decoded_bytes / ms / 1e6, "launches": plan.launches, "ok": ok}), flush=True)
