"""Where the time goes for small-vertex streams (debug-counter build: MOB200_LIB=.../dbg.so python tools/diag_c3.py kind [block])"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run

kind = sys.argv[1] if len(sys.argv) > 1 else "oct8"
block = len(sys.argv) > 2 and sys.argv[2] == "block"
total = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 24
seg = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 16
w = workloads.c3(kind, count=total, seg=seg) if kind in workloads.C3_KINDS else workloads.c2(total=total, seg=seg, keep_source=False)
P = loader.port()
sc = [P.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))[1] for i in range(w.n)] if block else None
outs, status, plan, guard = device_run(w, runs=0 if block else 3, sidecars=sc, block_runs=3 if block else 0)
plan.debug_counters(reset=True)
stream = torch.cuda.current_stream().cuda_stream
runs = 5
for _ in range(runs):
    plan.run(stream, block_parallel=block)
torch.cuda.synchronize()
d = plan.debug_counters(reset=True)
ms = min(plan.timing_history(runs))
dt, pt, wt = max(1, d["decoder_total"]), max(1, d["producer_total"]), max(1, d["walker_total"])
print(json.dumps({"kind": kind, "block": block, "streams": w.n, "ms": ms, "status_ok": bool((status == 0).all()),
                  "decoder_wait_full_pct": 100 * d["decoder_wait_full"] / dt, "decoder_wait_carry_pct": 100 * d["decoder_wait_carry"] / dt, "decoder_wait_tile_pct": 100 * d["decoder_wait_tile"] / dt,
                  "producer_meta_pct": 100 * d["producer_meta"] / pt, "producer_wait_slot_pct": 100 * d["producer_wait_slot"] / pt, "producer_lookback_pct": 100 * d["producer_lookback"] / pt,
                  "producer_cycles_per_run": pt / runs, "decoder_cycles_per_run": dt / runs, "walker_cycles_per_run": wt / runs}))
