"""Serial walk vs block mode (block-offset sidecar) on C2 segmentations and C1a; device-resident, best of N.
python tools/bench_blockmode.py [total_vertices]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
res = []
cases = [("C2 monolithic", lambda: workloads.c2(total=total, seg=None)),
         ("C2 x 65536", lambda: workloads.c2(total=total, seg=1 << 16)),
         ("C2 x 4096", lambda: workloads.c2(total=total, seg=1 << 12)),
         ("C1a v1", lambda: workloads.c1a(version=1, level=2)),
         ("C1a v0", lambda: workloads.c1a(version=0, level=0))]
for name, make in cases:
    w = make()
    runs = 2 if w.n <= 4 else 6
    outs, status, plan, guard = device_run(w, runs=runs, block_runs=6)
    hist = plan.timing_history(runs + 6)
    serial, block = min(hist[:runs]), min(hist[runs + 1:])
    ok = bool((status == 0).all() and guard and np.array_equal(np.concatenate(outs), w.source))
    r = {"workload": w.name, "streams": w.n, "serial_ms": serial, "block_ms": block, "serial_GBps": w.decoded_bytes / serial / 1e6, "block_GBps": w.decoded_bytes / block / 1e6,
         "plan_create_ms": plan.create_ms, "parity_ok": ok}
    print(json.dumps(r), flush=True)
    res.append(r)
