"""CPU tests of the multi-GPU host logic: the stream -> rank partition, exercised with a real
world_size-2 gloo process group (no GPU: each rank decodes its shard with the CPU checker and the ranks
exchange only checksums -- the data path itself has no collective)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_properties():
    from meshoptimizer_b200.sharding import contiguous_shards, shard_streams, stream_cost
    rng = np.random.default_rng(1)
    costs = [stream_cost(int(e), int(c), 32) for e, c in zip(rng.integers(100, 5000, 1000), rng.integers(1, 5000, 1000))]
    for world in (1, 2, 3, 8):
        shards = shard_streams(costs, world)
        allidx = sorted(i for s in shards for i in s)
        assert allidx == list(range(1000))            # every stream exactly once
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)   # LPT bound
        # the library's own partition (used by mob200_decode_batch_multi_host) is the same one
        from meshoptimizer_b200.sharding import shard_streams_native
        assert shard_streams_native(costs, world) == shards
    assert [len(r) for r in contiguous_shards(10, 4)] == [3, 3, 2, 2]
    assert shard_streams([], 2) == [[], []]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from meshoptimizer_b200.sharding import shard_streams, stream_cost
    from oracle import loader

    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    meta = data["codec_meta"]
    costs = [stream_cost(data[f"codec_{i}_enc"].size, int(c), int(vs)) for i, (c, vs, _, _) in enumerate(meta)]
    mine = shard_streams(costs, world)[rank]
    port_lib = loader.port()
    ok = 1
    checksum = np.zeros(len(meta), dtype=np.int64)
    for i in mine:
        c, vs = int(meta[i][0]), int(meta[i][1])
        rc, out = port_lib.decode_vertex_buffer(c, vs, data[f"codec_{i}_enc"])
        ok &= int(rc == 0 and np.array_equal(out, data[f"codec_{i}_dec"]))
        checksum[i] = int(out.astype(np.int64).sum()) + 1
    t = torch.from_numpy(checksum)
    dist.all_reduce(t)  # bookkeeping only: proves the shards are disjoint and complete
    flag = torch.tensor([ok])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    want = np.array([int(data[f"codec_{i}_dec"].astype(np.int64).sum()) + 1 for i in range(len(meta))])
    q.put((rank, bool(flag.item()), bool(np.array_equal(t.numpy(), want))))
    dist.destroy_process_group()


def test_two_rank_sharded_decode_gloo():
    import torch.multiprocessing as mp
    from oracle import loader
    loader.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, complete in results:
        assert ok, f"rank {rank}: a stream decoded incorrectly"
        assert complete, f"rank {rank}: shards were not disjoint and complete"
