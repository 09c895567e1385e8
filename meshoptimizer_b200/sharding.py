"""Stream -> GPU assignment for multi-GPU decode.  Streams are independent (the only state of the codec
lives inside a stream, reference src/vertexcodec.cpp:1846-1866), so sharding needs no data-path
collective: every rank decodes its own subset with its own context."""
from __future__ import annotations

from typing import List, Sequence


def stream_cost(encoded_bytes: int, vertex_count: int, vertex_size: int) -> int:
    """algorithmic bytes of one stream: encoded bytes read once + decoded bytes written once"""
    return int(encoded_bytes) + int(vertex_count) * int(vertex_size)


def shard_streams(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment: returns, per rank, the indices of its streams
    (ascending).  Deterministic, so every rank computes the same partition without communication."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    load = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(costs[i])
    for s in shards:
        s.sort()
    return shards


def contiguous_shards(n_streams: int, world_size: int) -> List[range]:
    """equal-count contiguous ranges (uniform streams)"""
    base, extra = divmod(n_streams, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append(range(start, start + cnt))
        start += cnt
    return out


def shard_streams_native(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """the same partition computed by the library (``mob200_shard_streams``, used by ``mob200_decode_batch_multi_host``)"""
    import ctypes

    from . import lib

    n = len(costs)
    arr = (ctypes.c_size_t * max(n, 1))(*[int(c) for c in costs])
    rank_of = (ctypes.c_int * max(n, 1))()
    rc = lib().mob200_shard_streams(arr, n, world_size, rank_of)
    if rc != 0:
        raise ValueError("mob200_shard_streams failed")
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in range(n):
        shards[rank_of[i]].append(i)
    return shards
