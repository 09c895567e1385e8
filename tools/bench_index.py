"""Secondary measurement: index-stream decode on the device (the row next to the hot path).  Many meshlet- and
mesh-sized triangle lists, reference-encoded, decoded by ONE launch of the thread-per-stream kernel with device-
resident buffers (CUDA events, best of N), next to the reference decoder on all host threads.  Every output is
verified against the reference outside the timed region.  Writes one JSON object per workload."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
from oracle import loader
from tests.index_cases import grid_triangles

R = loader.ref()
dev = torch.device("cuda:0")
ctx = mb.default_context()
res = []
for label, side, n_streams in (("meshlet-sized (126 triangles)", None, 200_000), ("32x32 grid patches (2048 triangles)", 32, 20_000), ("256x256 grid meshes (131072 triangles)", 256, 256)):
    if side is None:
        base = grid_triangles(8)[: 126 * 3]
    else:
        base = grid_triangles(side)
    rng = np.random.default_rng(3)
    variants = []
    for k in range(8):  # a few distinct streams, repeated: content differs, statistics are alike
        tri = base.reshape(-1, 3)
        variants.append(R.encode_index("triangles", rng.permutation(tri)[: tri.shape[0]].reshape(-1) if k else base, int(base.max()) + 1, 1))
    count = base.size
    pitch = (max(v.size for v in variants) + 15) & ~15
    blob = np.zeros(n_streams * pitch + 16, np.uint8)
    sizes = np.zeros(n_streams, np.int64)
    for i in range(n_streams):
        v = variants[i % 8]
        blob[i * pitch : i * pitch + v.size] = v
        sizes[i] = v.size
    d_src = torch.from_numpy(blob).to(dev)
    d_dst = torch.zeros(n_streams * count, dtype=torch.int16, device=dev)
    arr = (mb.IndexStream * n_streams)()
    for i in range(n_streams):
        arr[i].src = d_src.data_ptr() + i * pitch
        arr[i].src_size = int(sizes[i])
        arr[i].dst = d_dst.data_ptr() + 2 * i * count
        arr[i].index_count = count
        arr[i].index_size = 2
        arr[i].kind = mb.INDEX_TRIANGLES
    best = 1e9
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = mb.lib().mob200_decode_index_batch_device(ctx.handle, arr, n_streams, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        if it:
            best = min(best, e0.elapsed_time(e1))
    got = d_dst.cpu().numpy().view(np.uint16).reshape(n_streams, count)
    ok = True
    for k in range(8):
        rc, want = R.decode_index("triangles", count, 2, variants[k])
        ok = ok and rc == 0 and bool((got[k::8] == want[None, :]).all())
    streams = [(variants[i % 8], count, 2, 16) for i in range(min(n_streams, 50_000))]
    threads = R.hw_threads()
    cpu_s, _, _, st = R.decode_batch_mt(streams, threads, 3)
    assert all(s == 0 for s in st)
    decoded = n_streams * count * 2
    res.append({"workload": f"{n_streams} triangle lists, {label}, 16-bit indices, v1", "streams": n_streams, "decoded_MB": decoded / 1e6,
                "encoded_MB": float(sizes.sum()) / 1e6, "best_ms_incl_descriptor_upload_and_status": best,
                "triangles_per_second": n_streams * (count // 3) / (best * 1e-3), "decoded_GBps": decoded / best / 1e6,
                "cpu_reference_decoded_GBps": len(streams) * count * 2 / cpu_s / 1e9, "cpu_threads": threads, "parity_ok": ok})
print(json.dumps(res, indent=1))
