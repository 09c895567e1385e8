"""Seeded meshlets shared by the CPU (oracle) and GPU tests: (vertex references u32, triangles u8[n, 3])."""
from __future__ import annotations

import numpy as np


def _strip(nv: int, nt: int):
    """a triangle strip / fan mix over nv local vertices: what a clusterizer typically emits"""
    tris = []
    a, b, c = 0, 1, 2
    used = 3
    while len(tris) < nt:
        tris.append((a, b, c))
        if used < nv:
            a, b, c = b, c, used
            used += 1
        else:
            a, b, c = (a + 3) % nv, (b + 5) % nv, (c + 7) % nv
            if len({a, b, c}) < 3:
                a, b, c = 0, 1, 2
    return np.array(tris[:nt], np.uint8)


def meshlets(seed: int = 11):
    rng = np.random.default_rng(seed)
    yield "strip_64_124", np.arange(1000, 1064, dtype=np.uint32), _strip(64, 124)
    yield "strip_max", np.arange(7, 7 + 256, dtype=np.uint32) * 3, _strip(256, 256)
    yield "random_refs", rng.integers(0, 1 << 24, 96).astype(np.uint32), _strip(96, 180)
    yield "random_tris", np.sort(rng.integers(0, 50000, 128)).astype(np.uint32), rng.integers(0, 128, (200, 3)).astype(np.uint8)
    yield "wide_refs", rng.integers(0, 1 << 32, 33, dtype=np.uint64).astype(np.uint32), _strip(33, 40)
    yield "one", np.array([5, 6, 9], np.uint32), np.array([[0, 1, 2]], np.uint8)
    yield "odd_counts", np.arange(13, dtype=np.uint32) * 1000, _strip(13, 17)
    yield "triangles_only", np.zeros(0, np.uint32), _strip(40, 77)
    yield "empty", np.zeros(0, np.uint32), np.zeros((0, 3), np.uint8)


def corruptions(enc: np.ndarray, seed: int, n_random: int = 16):
    rng = np.random.default_rng(seed)
    for cut in (0, 1, enc.size // 2, max(0, enc.size - 1)):
        yield enc[:cut]
    yield np.concatenate([np.zeros(1, np.uint8), enc])
    for _ in range(n_random):
        e = enc.copy()
        k = rng.integers(0, e.size, 2)
        e[k] = rng.integers(0, 256, 2)
        yield e
