"""Diagnostics: when do the walker and decoder roles finish inside the fused kernel (library built with
-DMOB200_DEBUG_ENDS: tools/ab_build.py ends=MOB200_DEBUG_ENDS; MOB200_LIB=... python tools/gpu_ends.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
import bench

verts = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
segment = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 12
wl = bench.build_workload(verts, segment, 2, 1, 0, os.cpu_count() or 1)
n = len(wl["offsets"])
dev = torch.device("cuda:0")
ctx = mb.Context(0)
blob = torch.from_numpy(wl["blob"]).to(dev)
out_lens = (wl["counts"] * np.uint64(32) + np.uint64(15)) & ~np.uint64(15)
out_offs = np.zeros(n, np.uint64)
np.cumsum(out_lens[:-1], out=out_offs[1:])
out = torch.empty(int(out_lens.sum()) + 64, dtype=torch.uint8, device=dev)
items = [(blob.data_ptr() + int(wl["offsets"][i]), int(wl["sizes"][i]), out.data_ptr() + int(out_offs[i]), int(wl["counts"][i]), 32, 0) for i in range(n)]
plan = mb.Plan(ctx, mb.make_streams(items))
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    plan.run(stream)
torch.cuda.synchronize()
for _ in range(3):
    plan.debug_counters(reset=True)
    plan.run(stream)
    torch.cuda.synchronize()
    d = plan.debug_counters(reset=True)
    t = plan.timing_history(1)[-1]
    mhz = 1965.0
    print(f"decoder failed polls (cumulative, warp level) {d['walker_hard_waits']} |", end=" ")
    print(f"kernel {t:.3f} ms | walkers: mean end {d['r13']/max(1,(n+31)//32)/mhz/1e3:.3f} ms, last end {d['r15']/mhz/1e3:.3f} ms | decoders: last end {d['r14']/mhz/1e3:.3f} ms (cycles at {mhz:.0f} MHz)")
