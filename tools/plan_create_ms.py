"""Host time of mob200_plan_create against the number of streams (three plans per shape, freed in between)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import meshoptimizer_b200 as mb
ctx = mb.default_context()
dev = torch.device('cuda:0')
blob = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
out = torch.zeros(1 << 31, dtype=torch.uint8, device=dev)
for n, count in [(1024, 65536), (16384, 4096), (16384, 4096), (1024, 65536), (65536, 1024)]:
    items = [(blob.data_ptr(), 4096, out.data_ptr() + i * count * 32, count, 32, 0) for i in range(n)]
    t0 = time.perf_counter(); streams = mb.make_streams(items); t1 = time.perf_counter()
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        p = mb.Plan(ctx, streams); res.append(round(p.create_ms, 2)); del p
    print(n, count, 'make_streams ms', round((t1 - t0) * 1e3, 1), 'create_ms', res)
