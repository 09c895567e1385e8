"""Block-mode (sidecar) decode of the C3 fused-filter configurations: kernel ms and GB/s per kind, outputs checked
against the C restatement.  python tools/quick_c3.py [total] [seg] [kinds,comma,separated]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else list(workloads.C3_KINDS)
serial = len(sys.argv) > 4 and sys.argv[4] == "serial"  # first decode (no sidecar) instead of block mode
P = loader.port()
res = {}
for kind in kinds:
    w = workloads.c3(kind, count=total, seg=seg)
    sc = [P.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))[1] for i in range(w.n)]
    outs, status, plan, guard = device_run(w, runs=2, sidecars=None) if serial else device_run(w, runs=0, sidecars=sc, block_runs=2)
    ok = bool((status == 0).all()) and guard
    for i in range(0, w.n, max(1, w.n // 8)):
        rc, want = P.decode_vertex_buffer(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))
        want = P.decode_filter(w.meta["filter_name"], want, int(w.counts[i]), int(w.vertex_sizes[i]))
        ok = ok and rc == 0 and bool((outs[i] == want[: outs[i].size]).all())
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(8):
        plan.run(stream, block_parallel=not serial)
    torch.cuda.synchronize()
    ms = min(plan.timing_history(8))
    res[kind] = {"ms": round(ms, 4), "GBps": round(w.out_bytes() / ms / 1e6, 1), "ok": ok, "streams": w.n}
    del plan
print(json.dumps({"serial": serial, "rounds": os.environ.get("MOB200_ROUNDS", "auto"), "run_major": os.environ.get("MOB200_RUN_MAJOR", "1"), "total": total, "seg": seg, "c3_block_mode": res}))
