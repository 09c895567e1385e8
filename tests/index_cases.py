"""Seeded index-stream test cases shared by the CPU (oracle) and GPU tests."""
from __future__ import annotations

import numpy as np


def grid_triangles(side: int) -> np.ndarray:
    idx = []
    for y in range(side):
        for x in range(side):
            a = y * (side + 1) + x
            idx += [a, a + 1, a + side + 1, a + 1, a + side + 2, a + side + 1]
    return np.array(idx, np.uint32)


def index_sets(seed: int = 7):
    """(name, indices u32, vertex_count): regular, shuffled, random, restart-heavy and tiny inputs"""
    rng = np.random.default_rng(seed)
    g = grid_triangles(24)
    yield "grid", g, 625
    yield "grid_shuffled_triangles", rng.permutation(g.reshape(-1, 3)).reshape(-1), 625
    yield "random", rng.integers(0, 70000, 3 * 700).astype(np.uint32), 70000
    yield "strips", np.repeat(np.arange(0, 900, dtype=np.uint32), 3)[: 3 * 800] + np.tile(np.array([0, 1, 2], np.uint32), 800), 1000
    yield "restarts", np.tile(np.array([0, 1, 2, 2, 1, 3, 0, 1, 2, 2, 1, 5, 2, 1, 4], np.uint32), 40), 6
    yield "one_triangle", np.array([5, 6, 7], np.uint32), 8
    yield "empty", np.zeros(0, np.uint32), 1
    yield "wide", rng.integers(0, 1 << 29, 3 * 200).astype(np.uint32), 1 << 29


def corruptions(enc: np.ndarray, seed: int, n_random: int = 24):
    """truncated and byte-flipped variants of an encoded stream"""
    rng = np.random.default_rng(seed)
    for cut in (0, 1, 5, enc.size // 2, max(0, enc.size - 17), max(0, enc.size - 3), max(0, enc.size - 1)):
        yield enc[:cut]
    yield np.concatenate([enc, np.zeros(1, np.uint8)])
    for _ in range(n_random):
        e = enc.copy()
        k = rng.integers(0, e.size, 3)
        e[k] = rng.integers(0, 256, 3)
        yield e
