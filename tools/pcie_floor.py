"""The copy floor of the host-pointer path: every rank moves the bytes of one bench step between pinned host memory and
its GPU -- 0.91 GB in, 2.15 GB out, both directions at once, no kernels -- and the ranks do it at the same time.
    python tools/pcie_floor.py                                  (one GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/pcie_floor.py     (N GPUs of one box, NUMA-bound like bench.py)
Prints one JSON line: ms per step (max over ranks) and the decoded-equivalent GB/s of the whole box."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
numa = bench.numa_bind(local) if world > 1 else "single GPU: not bound"
if world > 1:
    import torch.distributed as dist
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n_out, n_in = 2147483648, 910869458
d = torch.empty(n_out, dtype=torch.uint8, device=dev); h = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n_in, dtype=torch.uint8, device=dev); h2 = torch.empty(n_in, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
best = 1e9
for it in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
    if it:
        best = min(best, dt)
if rank == 0:
    print(json.dumps({"n_gpus": world, "ms_per_step_max_over_ranks": best * 1e3, "decoded_equivalent_GBps_whole_box": world * n_out / best / 1e9,
                      "bytes_per_rank": {"h2d": n_in, "d2h": n_out}, "numa_rank0": numa}))
if world > 1:
    dist.destroy_process_group()
