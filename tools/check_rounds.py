"""Diagnostics: the rounds form forced on (MOB200_ROUNDS=1) over the few-long-stream C3 configurations, repeated; every byte compared."""
import os, sys
os.environ["MOB200_ROUNDS"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import workloads
from tests.gpu_util import device_run

bad = 0
for kind in ("oct8", "exp16", "oct12", "exp15"):
    w = workloads.c3(kind, count=1 << 24, seg=1 << 14, version=1, level=2)
    want = workloads.expected_outputs(w)
    stride = int(w.vertex_sizes[0])
    for rep in range(4):
        outs, status, plan, guard = device_run(w, runs=10)
        ok = bool((status == 0).all() and guard)
        for a, b in zip(outs, want):
            if stride == 4:
                d = np.abs(a.astype(np.int16) - b.astype(np.int16)); ok = ok and int(np.minimum(d, 256 - d).max()) <= 1
            else:
                ok = ok and np.array_equal(a, b)
        bad += not ok
        print(kind, rep, "ok" if ok else "MISMATCH", flush=True)
print("failures:", bad)
