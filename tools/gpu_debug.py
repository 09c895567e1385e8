"""First-contact GPU script: progressively harder cases with diagnostics (not a test)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run, first_mismatch

print(torch.cuda.get_device_name(0), mb.version(), "ref:", loader.have_ref(), flush=True)
kat = json.load(open(os.path.join(ROOT, "tests/golden/kat.json")))
for k in kat["codec"]:
    rc, out = mb.decode_vertex_buffer_rc(k["count"], k["size"], bytes.fromhex(k["input"]))
    ok = rc == 0 and out.tobytes().hex() == k["expected"]
    print("KAT", k["name"], "rc", rc, "OK" if ok else "MISMATCH", flush=True)
    if not ok:
        print("  got ", out.tobytes().hex()); print("  want", k["expected"])
R = loader.ref()
rng = np.random.default_rng(5)
for vs, count, ver, lvl in ((4, 16, 0, 0), (16, 256, 0, 0), (32, 256, 1, 2), (32, 1000, 1, 2), (32, 5000, 1, 3), (12, 777, 1, 2), (48, 500, 1, 2), (256, 100, 1, 2), (8, 3000, 0, 0)):
    v = np.cumsum(rng.integers(-9, 10, (count, vs)), axis=0).astype(np.uint8)
    enc = R.encode_vertex_buffer(v, count, vs, lvl, ver)
    rc, out = mb.decode_vertex_buffer_rc(count, vs, enc)
    print("RT vs", vs, "count", count, "v", ver, "L", lvl, "rc", rc, "mismatch", first_mismatch(out, v.reshape(-1)), flush=True)
for args in (dict(total=1 << 16, seg=1 << 12), dict(total=1 << 20, seg=1 << 14), dict(total=1 << 20, seg=None)):
    w = workloads.c2(**args)
    t = time.time()
    outs, status, plan, guard = device_run(w, runs=3)
    got = np.concatenate(outs)
    print(w.name, "status0", int((status == 0).sum()), "/", w.n, "guard", guard, "mismatch", first_mismatch(got, w.source), plan.last_timing(), flush=True)
