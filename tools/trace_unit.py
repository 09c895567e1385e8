"""Timeline of ONE decode unit (producer warp + four decoder warps) for a block-mode run of a C3 configuration.
Needs the trace build:  python tools/ab_build.py trace=MOB200_TRACE
    MOB200_LIB=meshoptimizer_b200/lib/variants/trace.so python tools/trace_unit.py oct8 [total] [seg] [first_block] [blocks]"""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run

kind = sys.argv[1] if len(sys.argv) > 1 else "oct8"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
seg = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 16
first = int(sys.argv[4]) if len(sys.argv) > 4 else 32
nblk = int(sys.argv[5]) if len(sys.argv) > 5 else 16
w = workloads.c3(kind, count=total, seg=seg) if kind in workloads.C3_KINDS else workloads.c2(total=total, seg=seg, keep_source=False)
P = loader.port()
sc = [P.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))[1] for i in range(w.n)]
outs, status, plan, guard = device_run(w, runs=0, sidecars=sc, block_runs=3)
L = mb.lib()
L.mob200_plan_debug_trace.restype = ctypes.c_int
L.mob200_plan_debug_trace.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16384)()
L.mob200_plan_debug_trace(plan.handle, buf, 16384)  # clear
stream = torch.cuda.current_stream().cuda_stream
plan.run(stream, block_parallel=True)
torch.cuda.synchronize()
L.mob200_plan_debug_trace(plan.handle, buf, 16384)
n = min(int(buf[0]), 16383)
ev = sorted(((int(buf[k]) >> 16, (int(buf[k]) >> 8) & 255, int(buf[k]) & 255) for k in range(1, n + 1)))
t0 = ev[0][0] if ev else 0
names = {1: "P meta begin", 2: "P meta end", 3: "P stage begin", 7: "P slot free", 4: "P stage end", 5: "P carry begin", 6: "P carry end"}
dn = ["round begin", "all members landed", "unpacked (carry wait)", "carry got", "tile written", "barrier passed", "stored+released"]
print(json.dumps({"kind": kind, "streams": w.n, "events": n, "ms": plan.timing_history(1)[0], "span_us": (ev[-1][0] - t0) / 1e3 if ev else 0}))
for t, e, i in ev:
    if not (first <= i < first + nblk):
        continue
    name = names.get(e) or f"D{(e - 32) // 8} {dn[(e - 32) % 8]}"
    print(f"{(t - t0) / 1e3:9.2f} us  blk {i:3d}  {name}")
