"""Standalone decode filters on device memory (mob200_filter_device; the reference's own filter benchmark is
tools/codecbench.cpp:108-162): in-place, so the algorithmic traffic is 2 * count * stride bytes.
CUDA events, best of N, inputs from the reference filter ENCODERS; a sample of every output is checked against the
reference filter decoders.  Next to it: the reference filters on all host threads."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from concurrent.futures import ThreadPoolExecutor

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
R = loader.ref()
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
threads = R.hw_threads()
res = []
base = 1 << 20  # encoder input generated once per kind and tiled (the filters are element-wise)
for kind in workloads.C3_KINDS:
    fname, stride, enc = workloads.c3_encoded_elements(kind, base)
    enc = np.ascontiguousarray(enc).view(np.uint8).reshape(-1)
    tiled = np.tile(enc, count // base)
    d = torch.from_numpy(tiled).to(dev)
    want = tiled[: base * stride].copy()
    R._filters[fname](want.ctypes.data, base, stride)
    best = 1e9
    for it in range(6):
        d.copy_(torch.from_numpy(tiled).to(dev)) if it else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mb.filter_device(fname, d.data_ptr(), count, stride, stream)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    got = d[: base * stride].cpu().numpy()
    last = d[-base * stride :].cpu().numpy()
    if stride == 4 and fname in ("oct", "color"):
        dd = np.abs(got.astype(np.int16) - want.astype(np.int16)); ok = int(np.minimum(dd, 256 - dd).max()) <= 1 and np.array_equal(got, last)
    else:
        ok = np.array_equal(got, want) and np.array_equal(last, want)
    # reference on the host threads: each thread filters its own slice in place
    hbuf = tiled[: min(tiled.size, (1 << 24) * stride)].copy()
    n_el = hbuf.size // stride
    per = (n_el // threads) & ~3
    f = R._filters[fname]
    def work(t):
        f(hbuf.ctypes.data + t * per * stride, per, stride)
    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter(); list(ex.map(work, range(threads))); cpu_s = time.perf_counter() - t0
    traffic = 2 * count * stride
    res.append({"filter": kind, "stride": stride, "elements": count, "ms": best, "traffic_GBps": traffic / best / 1e6, "roofline_frac": traffic / best / 1e6 / peak,
                "parity_ok": bool(ok), "cpu_reference_GBps": per * threads * stride / cpu_s / 1e9, "cpu_threads": threads})
    print(json.dumps(res[-1]), flush=True)
    del d
