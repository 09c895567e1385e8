"""One process, all GPUs of the box: SURVEY.md Appendix D "C5" through mob200_decode_batch_multi_host (host pointers;
LPT partition by algorithmic bytes, one host thread + context per GPU, no collective).

    python tools/bench_multi.py [n_gpus] [c2b_streams_per_gpu] [c3_elements]

Workload: n_gpus x C2b streams (65536 vertices x 32 B, v1 level 2, with sidecars) + the C3 set (seven filtered kinds,
64 Ki-element segments, v1 level 2, fused filters).  Reports the aggregate decoded GB/s, the per-device wall times and
the box's pinned-copy floor for the same bytes (every GPU copying its shard in and out concurrently, no kernels)."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
import bench
from oracle import loader, workloads

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
per_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c3_elems = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 22
threads = os.cpu_count() or 1
L = mb.lib()

# ---- workload in pinned host memory ----------------------------------------------------------------------------------
parts = []  # (blob, offsets, sizes, counts, vs, filter, sidecars, expected list)
v = bench.gen_vertices(0, n_gpus * per_gpu * 65536, threads)
wl = bench.encode_workload(v, 32, 1 << 16, 2, 1, threads, True)
parts.append((wl["blob"], wl["offsets"], wl["sizes"], wl["counts"], 32, 0, wl["sidecars"], [v]))
P = loader.port()
for kind in workloads.C3_KINDS:
    w = workloads.c3(kind, count=c3_elems, seg=1 << 16)
    exp = workloads.expected_outputs(w)
    sc = [P.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))[1] for i in range(w.n)]
    parts.append((w.blob, w.offsets, w.sizes, w.counts, int(w.vertex_sizes[0]), int(w.filters[0]), sc, exp))

in_bytes = sum(int(p[0].size) + 64 for p in parts)
out_bytes = sum(int((p[3] * np.uint64(p[4])).sum()) + 16 * len(p[3]) for p in parts)
h_in = torch.empty(in_bytes, dtype=torch.uint8).pin_memory()
h_out = torch.zeros(out_bytes, dtype=torch.uint8).pin_memory()
items, sidecars, checks = [], [], []
ci, co = 0, 0
for blob, offs, sizes, counts, vs, filt, sc, exp in parts:
    h_in[ci : ci + blob.size] = torch.from_numpy(blob)
    e = np.concatenate(exp)
    pos = 0
    for i in range(len(offs)):
        nb = int(counts[i]) * vs
        items.append((h_in.data_ptr() + ci + int(offs[i]), int(sizes[i]), h_out.data_ptr() + co, int(counts[i]), vs, filt))
        sidecars.append(sc[i])
        checks.append((co, e[pos : pos + nb], vs, filt))
        pos += nb
        co += (nb + 15) & ~15
    ci += (blob.size + 63) & ~63
arr = mb.make_streams(items)
n = len(items)
side = (ctypes.c_void_p * n)()
for i, s in enumerate(sidecars):
    side[i] = s.ctypes.data
devs = (ctypes.c_int * n_gpus)(*range(n_gpus))
dms = (ctypes.c_float * n_gpus)()
decoded = sum(it[3] * it[4] for it in items)
encoded = sum(it[1] for it in items)

def call():
    rc = L.mob200_decode_batch_multi_host(devs, n_gpus, arr, n, side, dms)
    assert rc == 0, rc

call()  # warm-up: contexts, arenas
ok = True
out_np = h_out.numpy()
for co, e, vs, filt in checks:
    got = out_np[co : co + e.size]
    if vs == 4 and filt in (1, 4):
        d = np.abs(got.astype(np.int16) - e.astype(np.int16)); ok = ok and int(np.minimum(d, 256 - d).max(initial=0)) <= 1
    else:
        ok = ok and np.array_equal(got, e)
times = []
for _ in range(3):
    t0 = time.perf_counter(); call(); times.append(time.perf_counter() - t0)
best = min(times)

# ---- pinned-copy floor: every GPU moves the bytes of its shard in and out concurrently, no kernels ----------------------
costs = (ctypes.c_size_t * n)(*[it[1] + it[3] * it[4] for it in items])
rank_of = (ctypes.c_int * n)()
L.mob200_shard_streams(costs, n, n_gpus, rank_of)
share_in = [sum(items[i][1] for i in range(n) if rank_of[i] == d) for d in range(n_gpus)]
share_out = [sum(items[i][3] * items[i][4] for i in range(n) if rank_of[i] == d) for d in range(n_gpus)]
bufs = []
oi = oo = 0
for d in range(n_gpus):
    with torch.cuda.device(d):
        bufs.append((torch.empty(share_in[d], dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(share_out[d], dtype=torch.uint8, device=f"cuda:{d}"),
                     torch.cuda.Stream(device=d), torch.cuda.Stream(device=d), oi, oo))
    oi += share_in[d]; oo += share_out[d]
def floor_pass():
    for d, (di, do, s1, s2, a, b) in enumerate(bufs):
        with torch.cuda.stream(s1):
            di.copy_(h_in[a : a + di.numel()], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[b : b + do.numel()].copy_(do, non_blocking=True)
    for d in range(n_gpus):
        torch.cuda.synchronize(d)
floor_pass()
ft = []
for _ in range(3):
    t0 = time.perf_counter(); floor_pass(); ft.append(time.perf_counter() - t0)
floor = min(ft)
print(json.dumps({"n_gpus": n_gpus, "streams": n, "decoded_GB": decoded / 1e9, "encoded_GB": encoded / 1e9, "parity_ok": bool(ok),
                  "e2e_ms": best * 1e3, "e2e_decoded_GBps": decoded / best / 1e9, "device_ms": [float(x) for x in dms],
                  "shard_decoded_GB": [s / 1e9 for s in share_out], "copy_floor_ms": floor * 1e3, "copy_floor_decoded_GBps": decoded / floor / 1e9,
                  "e2e_over_floor": floor / best, "api": "mob200_decode_batch_multi_host (pinned host buffers, sidecars, LPT by algorithmic bytes)"}))
