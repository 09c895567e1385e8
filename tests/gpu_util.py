"""Helpers for the -m gpu tests: run workloads through the C ABI with device-resident buffers."""
from __future__ import annotations

import numpy as np


def device_run(w, ctx=None, dst_shift: int = 0, src_shift: int = 0, runs: int = 1, sidecars=None, block_runs: int = 0, blob_override=None):
    """Decode a oracle.workloads.Workload with mob200_plan_* on cuda:0.
    runs serial-walk runs, then block_runs block-parallel runs (from `sidecars` or the offsets the serial runs left).
    Returns (list of per-stream uint8 outputs, status array, plan, guard bytes untouched)."""
    import torch
    import meshoptimizer_b200 as mb

    ctx = ctx or mb.default_context()
    dev = torch.device("cuda:0")
    blob = torch.zeros(w.blob.size + 32 + src_shift, dtype=torch.uint8, device=dev)
    blob[src_shift : src_shift + w.blob.size] = torch.from_numpy(w.blob if blob_override is None else blob_override).to(dev)
    out_offs = w.out_offsets()
    out = torch.full((w.out_bytes() + 64 + dst_shift,), 0xCD, dtype=torch.uint8, device=dev)
    items = []
    for i in range(w.n):
        items.append((blob.data_ptr() + src_shift + int(w.offsets[i]), int(w.sizes[i]), out.data_ptr() + dst_shift + int(out_offs[i]),
                      int(w.counts[i]), int(w.vertex_sizes[i]), int(w.filters[i])))
    plan = mb.Plan(ctx, mb.make_streams(items), sidecars=sidecars)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(runs):
        plan.run(stream)
    if block_runs:
        if runs:
            out.fill_(0xCD)  # the block-mode runs must produce every byte themselves
        for _ in range(block_runs):
            plan.run(stream, block_parallel=True)
    status = plan.status(stream)
    host = out.cpu().numpy()
    outs = []
    for i in range(w.n):
        o = dst_shift + int(out_offs[i])
        outs.append(host[o : o + int(w.counts[i]) * int(w.vertex_sizes[i])])
    # bytes between streams must be untouched (no write slack may be assumed)
    guard_ok = True
    for i in range(w.n):
        end = dst_shift + int(out_offs[i]) + int(w.counts[i]) * int(w.vertex_sizes[i])
        nxt = dst_shift + int(out_offs[i + 1]) if i + 1 < w.n else host.size
        if not (host[end:nxt] == 0xCD).all():
            guard_ok = False
            break
    return outs, status, plan, guard_ok


def first_mismatch(a: np.ndarray, b: np.ndarray):
    if a.size != b.size:
        return ("size", a.size, b.size)
    d = np.nonzero(a != b)[0]
    if d.size == 0:
        return None
    i = int(d[0])
    return (i, int(a[i]), int(b[i]), int(d.size))
