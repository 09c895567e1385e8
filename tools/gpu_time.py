"""Quick device-resident timing of C2-style workloads (not the bench; diagnostics only)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import workloads
from tests.gpu_util import device_run, first_mismatch

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
for seg in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4096", "65536"])]:
    w = workloads.c2(total=total, seg=seg)
    outs, status, plan, guard = device_run(w, runs=3)
    ok = np.array_equal(np.concatenate(outs), w.source)
    t = plan.last_timing()
    alg = w.encoded_bytes + w.decoded_bytes
    print(f"{w.name}: ok={ok} walk {t['walk_ms']:.3f} ms decode {t['decode_ms']:.3f} ms | decode-only {alg/t['decode_ms']/1e6:.0f} GB/s traffic, "
          f"{w.decoded_bytes/t['decode_ms']/1e6:.0f} GB/s decoded | total {w.decoded_bytes/t['total_ms']/1e6:.0f} GB/s decoded", flush=True)
