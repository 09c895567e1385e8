"""Synthetic workloads C1-C5 (SURVEY.md Appendix D / BASELINE.md section 4).  TEST / BENCH INFRASTRUCTURE.

Every stream is produced by the UNMODIFIED reference encoder (oracle/_ref, built from
/root/reference by oracle/Makefile), as north_star requires.  Imported by tests/, bench.py and
__graft_entry__.smoke() only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import loader


@dataclass
class Workload:
    name: str
    blob: np.ndarray            # all encoded streams, each starting on a 16-byte boundary
    offsets: np.ndarray         # u64[n]
    sizes: np.ndarray           # u64[n]
    counts: np.ndarray          # u64[n] vertices per stream
    vertex_sizes: np.ndarray    # u32[n]
    filters: np.ndarray         # i32[n]
    source: Optional[np.ndarray] = None   # original vertex bytes when all streams are slices of one array (no filter)
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.offsets.size)

    @property
    def decoded_bytes(self) -> int:
        return int((self.counts * self.vertex_sizes).sum())

    @property
    def encoded_bytes(self) -> int:
        return int(self.sizes.sum())

    def stream(self, i: int) -> np.ndarray:
        o, s = int(self.offsets[i]), int(self.sizes[i])
        return self.blob[o : o + s]

    def out_offsets(self) -> np.ndarray:
        """16-byte aligned output offsets, one region per stream."""
        lens = (self.counts * self.vertex_sizes + 15) & ~np.uint64(15)
        offs = np.zeros(self.n, dtype=np.uint64)
        np.cumsum(lens[:-1], out=offs[1:])
        return offs

    def out_bytes(self) -> int:
        lens = (self.counts * self.vertex_sizes + 15) & ~np.uint64(15)
        return int(lens.sum())

    def harness_streams(self):
        return [(self.stream(i), int(self.counts[i]), int(self.vertex_sizes[i]), int(self.filters[i])) for i in range(self.n)]


def _segments(total: int, seg: int):
    firsts = np.arange(0, total, seg, dtype=np.uint64)
    counts = np.minimum(np.uint64(seg), np.uint64(total) - firsts).astype(np.uint64)
    return firsts, counts


def from_vertices(name: str, vertices: np.ndarray, vertex_size: int, seg: Optional[int], level: int, version: int, filter_id: int = 0, keep_source: bool = True) -> Workload:
    R = loader.ref()
    v = np.ascontiguousarray(vertices).view(np.uint8).reshape(-1)
    total = v.size // vertex_size
    firsts, counts = _segments(total, seg or max(total, 1))
    if total == 0:
        firsts, counts = np.zeros(1, np.uint64), np.zeros(1, np.uint64)
    blob, offs, sizes = R.encode_segments(v, vertex_size, firsts, counts, level, version)
    n = firsts.size
    return Workload(name, blob, offs, sizes, counts, np.full(n, vertex_size, np.uint32), np.full(n, filter_id, np.int32),
                    source=v if keep_source else None, meta={"level": level, "version": version, "segment": seg})


# ---- C1 ---------------------------------------------------------------------------------------------

def c1a(version: int = 1, level: int = 2, side: int = 1000) -> Workload:
    """codecbench grid: (side+1)^2 vertices x 32 bytes, cache+fetch reordered, one stream."""
    R = loader.ref()
    v = R.grid_reorder(R.gen_grid(side), side)
    return from_vertices(f"C1a grid {side+1}^2 x 32B v{version} L{level}", v, 32, None, level, version)


def c1b(version: int = 0, level: int = 2, count: int = 1 << 20) -> Workload:
    """js/benchmark.js stream: count vertices x 16 bytes, one stream."""
    v = loader.port().gen_js16(count)
    return from_vertices(f"C1b js16 {count} x 16B v{version} L{level}", v, 16, None, level, version)


# ---- C2 ---------------------------------------------------------------------------------------------

def c2(total: int = 1 << 26, seg: Optional[int] = 1 << 16, level: int = 2, version: int = 1, keep_source: bool = True) -> Workload:
    """32-byte vertices (grid words + counter + quantised float); seg=None -> one monolithic stream."""
    v = loader.port().gen_c2(0, total)
    tag = "monolithic" if not seg else f"{(total + seg - 1) // seg} x {seg}-vertex streams"
    return from_vertices(f"C2 {total} x 32B v{version} L{level}, {tag}", v, 32, seg, level, version, keep_source=keep_source)


# ---- C3: gltfpack-style filtered streams ---------------------------------------------------------------

def _fib_sphere(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    z = 1.0 - 2.0 * i / n
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = i * (np.pi * (3.0 - np.sqrt(5.0)))
    out = np.zeros((n, 4), dtype=np.float32)
    out[:, 0] = r * np.cos(phi)
    out[:, 1] = r * np.sin(phi)
    out[:, 2] = z
    return out


def c3_encoded_elements(kind: str, count: int) -> tuple:
    """Returns (filter name, stride, encoded element bytes) produced by the reference filter encoders,
    with the gltfpack default bit counts (gltf/gltfpack.cpp:1254-1260, gltf/stream.cpp:591-855)."""
    R = loader.ref()
    i = np.arange(count, dtype=np.float64)
    if kind in ("oct8", "oct12"):
        data = _fib_sphere(count)
        data[:, 3] = (np.arange(count) % 2) * 2.0 - 1.0
        stride, bits = (4, 8) if kind == "oct8" else (8, 12)
        return "oct", stride, R.encode_filter("oct", data, count, stride, bits)
    if kind == "quat12":
        a = i * 1e-4
        q = np.stack([np.sin(a), np.sin(2 * a), np.cos(3 * a), np.cos(a)], axis=1)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        return "quat", 8, R.encode_filter("quat", q.astype(np.float32), count, 8, 12)
    if kind in ("exp15", "exp16"):
        idx = np.arange(count, dtype=np.int64)
        h = ((idx * 2654435761) & 0xFFFF).astype(np.float64) / 65536.0
        p = np.stack([(idx % 4096) * 0.125 + 0.001 * h, ((idx // 4096) % 4096) * 0.125 + 0.001 * h, (idx >> 24) * 0.125 + 0.001 * h], axis=1)
        bits, mode = (15, 0) if kind == "exp15" else (16, 1)  # Separate / SharedVector
        return "exp", 12, R.encode_filter("exp", p.astype(np.float32), count, 12, bits, mode)
    if kind in ("color8", "color12"):
        a = i * 1e-3
        c = np.stack([0.5 + 0.5 * np.sin(a), 0.5 + 0.5 * np.sin(1.3 * a + 1), 0.5 + 0.5 * np.cos(0.7 * a), 0.5 + 0.5 * np.cos(0.1 * a)], axis=1)
        stride, bits = (4, 8) if kind == "color8" else (8, 12)
        return "color", stride, R.encode_filter("color", c.astype(np.float32), count, stride, bits)
    raise ValueError(kind)


C3_KINDS = ("oct8", "oct12", "quat12", "exp15", "exp16", "color8", "color12")


def c3(kind: str, count: int = 1 << 24, seg: int = 1 << 16, version: int = 1, level: int = 2) -> Workload:
    fname, stride, enc = c3_encoded_elements(kind, count)
    w = from_vertices(f"C3 {kind} {count} x {stride}B v{version} L{level} filter={fname}", enc, stride, seg, level, version,
                      filter_id=loader.FILTER_NAMES[fname], keep_source=True)
    w.meta["filter_name"] = fname
    return w


# ---- C4: many meshlet-sized streams ------------------------------------------------------------------------

def _murmur(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def c4(n_streams: int = 200_000, version: int = 1, level: int = 2) -> Workload:
    """n independent streams of 64..256 vertices, vertex sizes 12/16/32 round-robin."""
    R = loader.ref()
    P = loader.port()
    counts = (64 + (_murmur(np.arange(n_streams, dtype=np.uint32)) % np.uint32(193))).astype(np.uint64)
    vss = np.array([12, 16, 32], dtype=np.uint32)[np.arange(n_streams) % 3]
    pool = P.gen_c2(0, 1 << 20).view(np.uint8).reshape(-1)  # content pool, sliced per stream
    blobs, offs, sizes = [], np.zeros(n_streams, np.uint64), np.zeros(n_streams, np.uint64)
    cursor = 0
    for vs in (12, 16, 32):
        idx = np.nonzero(vss == vs)[0]
        if idx.size == 0:
            continue
        # stream i reads count_i vertices of vs bytes starting at byte offset (i*256*32) % pool
        lens = counts[idx] * np.uint64(vs)
        starts = (idx.astype(np.uint64) * np.uint64(256 * 32)) % np.uint64(pool.size - 256 * 32)
        flat = np.concatenate([pool[int(s) : int(s) + int(l)] for s, l in zip(starts, lens)])
        firsts = np.zeros(idx.size, np.uint64)
        np.cumsum(counts[idx][:-1], out=firsts[1:])
        b, o, s = R.encode_segments(flat, vs, firsts, counts[idx], level, version)
        blobs.append(b)
        offs[idx] = o + np.uint64(cursor)
        sizes[idx] = s
        cursor += b.size
    blob = np.concatenate(blobs) if blobs else np.zeros(16, np.uint8)
    return Workload(f"C4 {n_streams} streams of 64-256 vertices, vs 12/16/32, v{version} L{level}", blob, offs, sizes, counts, vss,
                    np.zeros(n_streams, np.int32), source=None, meta={"level": level, "version": version})


def merge(name: str, parts: List[Workload], interleave: bool = True) -> Workload:
    """One batch out of several workloads (mixed vertex sizes / filters); interleave=True deals the streams round-robin."""
    blobs, offs, cursor = [], [], 0
    for w in parts:
        pad = (-w.blob.size) % 16
        blobs.append(np.concatenate([w.blob, np.zeros(pad, np.uint8)]))
        offs.append(w.offsets + np.uint64(cursor))
        cursor += w.blob.size + pad
    cat = lambda f: np.concatenate([getattr(w, f) for w in parts])
    offsets, sizes, counts, vss, filters = np.concatenate(offs), cat("sizes"), cat("counts"), cat("vertex_sizes"), cat("filters")
    if interleave:
        ids = np.concatenate([np.arange(w.n, dtype=np.float64) / max(w.n, 1) for w in parts])
        order = np.argsort(ids, kind="stable")
        offsets, sizes, counts, vss, filters = offsets[order], sizes[order], counts[order], vss[order], filters[order]
    return Workload(name, np.concatenate(blobs), offsets, sizes, counts, vss, filters, source=None, meta={"parts": [w.name for w in parts]})


def expected_outputs(w: Workload, lib=None, threads: int = 0) -> List[np.ndarray]:
    """Decode (and filter) every stream with a CPU checker (reference when available, else the port)."""
    lib = lib or (loader.ref() if loader.have_ref() else loader.port())
    _, _, outs, status = lib.decode_batch_mt(w.harness_streams(), threads or lib.hw_threads(), 1)
    assert all(s == 0 for s in status), "CPU checker rejected a generated stream"
    return outs
