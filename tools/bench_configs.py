"""Secondary measurements on the other BASELINE.json configs (C1a/C1b single streams, C3 fused filters,
C4 many small streams, few-long-stream variants of C2).  Device-resident, CUDA events, best of N runs;
every output is verified against the CPU checker outside the timed region.  Writes one JSON object."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run

def measure(w, runs=10, check="source"):
    outs, status, plan, guard = device_run(w, runs=runs)
    hist = plan.timing_history(runs)
    best = min(h for h in hist[2:]) if len(hist) > 2 else hist[-1]
    ok = bool((status == 0).all() and guard)
    if check == "source":
        ok = ok and np.array_equal(np.concatenate(outs), w.source)
    else:
        want = workloads.expected_outputs(w)
        stride = int(w.vertex_sizes[0])
        for a, b in zip(outs, want):
            if stride == 4 and check in ("oct8", "color8"):
                d = np.abs(a.astype(np.int16) - b.astype(np.int16)); ok = ok and int(np.minimum(d, 256 - d).max()) <= 1
            else:
                ok = ok and np.array_equal(a, b)
    return {"workload": w.name, "streams": w.n, "decoded_MB": w.decoded_bytes / 1e6, "ratio": w.encoded_bytes / max(1, w.decoded_bytes),
            "best_ms": best, "decoded_GBps": w.decoded_bytes / best / 1e6, "traffic_GBps": (w.decoded_bytes + w.encoded_bytes) / best / 1e6, "parity_ok": ok}

res = []
res.append(measure(workloads.c1a(version=0, level=0)))
res.append(measure(workloads.c1a(version=1, level=2)))
res.append(measure(workloads.c1b(version=0)))
res.append(measure(workloads.c1b(version=1)))
for seg in (None, 1 << 16, 1 << 12, 1 << 8):
    res.append(measure(workloads.c2(total=1 << 24, seg=seg)))
res.append(measure(workloads.c2(total=1 << 24, seg=1 << 12, level=3)))
res.append(measure(workloads.c2(total=1 << 24, seg=1 << 12, level=0, version=0)))
for kind in workloads.C3_KINDS:
    for version in (0, 1):
        res.append(measure(workloads.c3(kind, count=1 << 24, seg=1 << 14, version=version, level=2 if version else 0), check=kind))
w4 = workloads.c4(200_000)
r = measure(w4, check="c4"); r["streams_per_second"] = w4.n / (r["best_ms"] * 1e-3); res.append(r)
w1 = workloads.c4(1)
r = measure(w1, runs=20, check="c4"); r["note"] = "latency of a single <=256-vertex stream (one launch)"; res.append(r)
print(json.dumps(res, indent=1))
