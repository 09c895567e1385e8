"""Quick device-resident timing of the two main shapes (for A/B of variant builds: MOB200_LIB=... python tools/quick.py).
   python tools/quick.py [verts] [steps]   -> one JSON line: c2b + sidecar (block mode) and 4096-vertex streams (serial walk)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import meshoptimizer_b200 as mb
import bench

verts = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
which = sys.argv[3].split(",") if len(sys.argv) > 3 else ["c2b_sidecar", "seg4096"]
threads = os.cpu_count() or 1
dev = torch.device("cuda:0")
ctx = mb.Context(0)
stream = torch.cuda.current_stream().cuda_stream
v = bench.gen_vertices(0, verts, threads)
expected = torch.from_numpy(v).to(dev)
out = torch.empty(v.size + 64, dtype=torch.uint8, device=dev)
res = {"lib": os.environ.get("MOB200_LIB", "default")}
for name in which:
    seg, block = bench.headline_shape(name)
    wl = bench.encode_workload(v, 32, seg, 2, 1, threads, block)
    dw = bench.DeviceWorkload(mb, ctx, dev, wl, out, block)
    ok = dw.parity(expected, stream)
    ms, kmean, kbest = bench.time_steps(dw, stream, steps, 3, torch.cuda.synchronize)
    res[name] = {"ms": round(ms, 4), "best": round(kbest, 4), "ok": ok}
    del dw
print(json.dumps(res))
