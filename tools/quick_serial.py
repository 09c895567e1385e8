"""first-decode (serial path) timings: C2b 1024 x 65536, monolithic 4 Mi, C1a v0/v1.  MOB200_LIB=... python tools/quick_serial.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import meshoptimizer_b200 as mb
import bench
from oracle import workloads
from tests.gpu_util import device_run
threads = os.cpu_count() or 1
dev = torch.device("cuda:0"); ctx = mb.Context(0); stream = torch.cuda.current_stream().cuda_stream
v = bench.gen_vertices(0, 1 << 26, threads)
expected = torch.from_numpy(v).to(dev); out = torch.empty(v.size + 64, dtype=torch.uint8, device=dev)
res = {"lib": os.path.basename(os.environ.get("MOB200_LIB", "default"))}
wl = bench.encode_workload(v, 32, 1 << 16, 2, 1, threads, False)
dw = bench.DeviceWorkload(mb, ctx, dev, wl, out, False)
ok = dw.parity(expected, stream)
ms, kmean, kbest = bench.time_steps(dw, stream, 5, 2, torch.cuda.synchronize)
res["c2b_serial_ms"] = round(kbest, 3); res["ok"] = ok
del dw, expected, out
for w in (workloads.c2(total=1 << 22, seg=None), workloads.c1a(version=1, level=2), workloads.c1a(version=0, level=0)):
    outs, status, plan, guard = device_run(w, runs=3)
    res[w.name[:14]] = round(min(plan.timing_history(3)), 2)
print(json.dumps(res))
