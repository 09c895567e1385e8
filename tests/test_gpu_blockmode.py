"""-m gpu: block mode -- every block walked by its own lane from a block-offset sidecar (include/meshopt_b200.h
section 2b; the serial loop it replaces: reference src/vertexcodec.cpp:1857-1866).

* a plan that has run once decodes again block-parallel, bit-exact (monolithic, segmented, filtered, tiny streams);
* exported sidecars equal the CPU checker's block offsets, and a fresh plan created WITH them decodes without
  ever running the serial walk;
* a sidecar is never trusted: a shifted entry, a sidecar of another stream, input bytes changed after the walk
  and random garbage are all rejected with MOB200_ERR_SIDECAR (no hang, no write outside the outputs).
"""
import numpy as np
import pytest

from oracle import loader, workloads
from tests.gpu_util import device_run, first_mismatch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import meshoptimizer_b200 as m

    m.lib()
    return m


def _expected(w):
    if w.source is not None and int(w.filters.max()) == 0:
        outs, o = [], 0
        for i in range(w.n):
            nbytes = int(w.counts[i]) * int(w.vertex_sizes[i])
            outs.append(w.source[o : o + nbytes])
            o += nbytes
        return outs
    return workloads.expected_outputs(w)


def _same(w, outs, want):
    for i, (a, b) in enumerate(zip(outs, want)):
        if int(w.vertex_sizes[i]) == 4 and int(w.filters[i]) in (1, 4):
            d = np.abs(a.astype(np.int16) - b.astype(np.int16))
            if int(np.minimum(d, 256 - d).max(initial=0)) > 1:
                return f"stream {i}: more than 1 LSB apart"
        else:
            m = first_mismatch(a, b)
            if m is not None:
                return f"stream {i}: {m}"
    return None


WORKLOADS = {
    "c2_mono_256k": lambda: workloads.c2(total=1 << 18, seg=None),
    "c2_mono_ragged": lambda: workloads.c2(total=100_003, seg=None),
    "c2_seg_4096": lambda: workloads.c2(total=1 << 19, seg=1 << 12),
    "c2_l3_xor": lambda: workloads.c2(total=1 << 17, seg=1 << 15, level=3),
    "c2_v0": lambda: workloads.c2(total=1 << 17, seg=None, level=0, version=0),
    "c1a_v0": lambda: workloads.c1a(version=0, level=0, side=200),
    "c1b_v1": lambda: workloads.c1b(version=1, count=1 << 16),
    "c3_oct8": lambda: workloads.c3("oct8", count=1 << 17, seg=1 << 15),
    "c3_quat12_v0": lambda: workloads.c3("quat12", count=70_000, seg=None, version=0, level=0),
    "c3_exp15": lambda: workloads.c3("exp15", count=1 << 16, seg=1 << 14),
    "c4": lambda: workloads.c4(3000),
}


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_block_mode_reuses_offsets(mb, name):
    w = WORKLOADS[name]()
    want = _expected(w)
    outs, status, plan, guard = device_run(w, runs=1, block_runs=2)
    assert plan.has_offsets
    assert (status == 0).all(), status[status != 0][:8]
    assert guard
    assert _same(w, outs, want) is None, _same(w, outs, want)


@pytest.mark.parametrize("name", ["c2_mono_256k", "c2_seg_4096", "c1a_v0", "c3_oct8", "c4"])
def test_sidecar_export_import(mb, name):
    w = WORKLOADS[name]()
    want = _expected(w)
    outs, status, plan, guard = device_run(w, runs=1)
    assert (status == 0).all()
    port = loader.port()
    sidecars = []
    for i in range(w.n):
        sc = plan.export_sidecar(i, int(w.counts[i]), int(w.vertex_sizes[i]))
        rc, off = port.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))
        assert rc == 0 and np.array_equal(sc, off), f"stream {i}: exported sidecar differs from the CPU walk"
        sidecars.append(sc.copy())
    del plan
    # a fresh plan that only ever runs in block mode
    outs2, status2, plan2, guard2 = device_run(w, runs=0, sidecars=sidecars, block_runs=1)
    assert (status2 == 0).all() and guard2
    assert _same(w, outs2, want) is None


def _sidecars(w):
    port = loader.port()
    out = []
    for i in range(w.n):
        rc, off = port.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))
        assert rc == 0
        out.append(off.copy())
    return out


def test_stale_sidecar_is_rejected(mb):
    w = workloads.c2(total=1 << 17, seg=1 << 14)  # 8 streams x 64 blocks
    want = _expected(w)
    sc = _sidecars(w)
    bad = [s.copy() for s in sc]
    bad[1][10] += 1             # one block start off by one
    bad[3] = sc[4].copy()       # another stream's table (same length, different data)
    bad[5][-1] -= 1             # the end of the last block
    bad[6][0] = 2               # block 0 must start at byte 1
    outs, status, plan, guard = device_run(w, runs=0, sidecars=bad, block_runs=1)
    assert guard
    assert list(status) == [0, mb.ERR_SIDECAR, 0, mb.ERR_SIDECAR, 0, mb.ERR_SIDECAR, mb.ERR_SIDECAR, 0], status
    for i in (0, 2, 4, 7):
        assert np.array_equal(outs[i], want[i])
    # decoding again without the sidecar gives the reference result for every stream
    outs, status, plan, guard = device_run(w, runs=1)
    assert (status == 0).all() and _same(w, outs, want) is None


def test_changed_input_is_caught(mb):
    """offsets kept from a walk of other bytes: the block that no longer ends where the table says is rejected"""
    import torch

    w = workloads.c2(total=1 << 16, seg=None)
    sc = _sidecars(w)
    blob = w.blob.copy()
    # clear the selector byte of a bit-packed channel header in block 3: its groups change width, the block shrinks
    o = int(w.offsets[0]) + int(sc[0][3]) + 8  # 8 control bytes, then channel 0's header
    assert blob[o] != 0
    blob[o] = 0
    rc, _ = loader.ref().decode_vertex_buffer(int(w.counts[0]), 32, blob[int(w.offsets[0]) : int(w.offsets[0]) + int(w.sizes[0])]) if loader.have_ref() else (-1, None)
    outs, status, plan, guard = device_run(w, runs=0, sidecars=sc, block_runs=1, blob_override=blob)
    assert guard and status[0] == mb.ERR_SIDECAR
    if loader.have_ref():
        assert rc != 0  # (the reference rejects the modified stream as well)


def test_garbage_sidecars_are_safe(mb):
    rng = np.random.default_rng(5)
    w = workloads.c2(total=1 << 16, seg=1 << 12)
    sc = _sidecars(w)
    bad = []
    for i, s in enumerate(sc):
        g = s.copy()
        kind = i % 4
        if kind == 0:
            g[:] = rng.integers(0, 1 << 32, size=g.size, dtype=np.uint64).astype(np.uint32)
        elif kind == 1:
            g[:] = rng.integers(0, int(w.sizes[i]) + 64, size=g.size).astype(np.uint32)
        elif kind == 2:
            g[:] = np.sort(rng.integers(1, int(w.sizes[i]), size=g.size)).astype(np.uint32)
            g[0] = 1
        else:
            g[1:-1] = g[1:-1][::-1]
        bad.append(g)
    outs, status, plan, guard = device_run(w, runs=0, sidecars=bad, block_runs=2)
    assert guard
    assert (status == mb.ERR_SIDECAR).all(), status


def test_block_mode_needs_offsets(mb):
    w = workloads.c2(total=4096, seg=None)
    import torch

    with pytest.raises(RuntimeError):
        device_run(w, runs=0, block_runs=1)


def test_block_mode_error_streams(mb, checker):
    """framing errors keep their reference codes in block mode; empty streams are checked too"""
    w = workloads.c2(total=1 << 14, seg=1 << 12)
    sc = _sidecars(w)
    blob = w.blob.copy()
    blob[int(w.offsets[1])] = 0x55       # bad magic -> -1
    blob[int(w.offsets[2])] = 0xA2       # version 2 -> -1
    outs, status, plan, guard = device_run(w, runs=0, sidecars=sc, block_runs=1, blob_override=blob)
    assert guard and list(status) == [0, -1, -1, 0]


@pytest.mark.parametrize("vs,total,seg,version,level", [(32, 300_000, 65_536, 1, 2), (32, 200_000, 0, 1, 2), (16, 150_000, 4096, 0, 0), (12, 90_000, 0, 1, 3), (8, 70_001, 10_000, 1, 2)])
def test_segmenter_streams_and_sidecars_decode_on_the_device(mb, vs, total, seg, version, level):
    """mob200_encode_segments (host) -> streams + sidecars -> block-mode decode, never running the serial walk"""
    rng = np.random.default_rng(9)
    words = np.cumsum(rng.integers(-200, 200, (total, vs // 4)), axis=0).astype(np.uint32)
    v = words.view(np.uint8).reshape(-1)
    blob, segs, side = mb.encode_segments(v, total, vs, seg, level, version, threads=0)
    n = len(segs)
    w = workloads.Workload("segmenter", np.concatenate([blob, np.zeros(32, np.uint8)]),
                           np.array([s.offset for s in segs], np.uint64), np.array([s.size for s in segs], np.uint64),
                           np.array([s.vertex_count for s in segs], np.uint64), np.full(n, vs, np.uint32), np.zeros(n, np.int32), source=v)
    sidecars = [side[s.sidecar_offset : s.sidecar_offset + s.sidecar_entries].copy() for s in segs]
    outs, status, plan, guard = device_run(w, runs=0, sidecars=sidecars, block_runs=1)
    assert (status == 0).all() and guard
    assert np.array_equal(np.concatenate(outs), v)


# ---- block mode + rounds form: run-major decode order, carries chained inside a round by the decoder warps ---------------

@pytest.fixture
def rounds_ctx(mb, monkeypatch):
    """a context with the rounds form forced on (the heuristic wants >= 8 blocks per decode unit)"""
    monkeypatch.setenv("MOB200_ROUNDS", "1")
    ctx = mb.Context(-1)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("kind,count,seg,version,level", [
    ("oct8", 800_000, None, 1, 2),       # one stream, 3125 four-byte blocks: every round is a run of four blocks of it
    ("quat12", 800_003, None, 1, 2),     # 8-byte vertices, ragged last block
    ("exp15", 900_000, None, 1, 2),      # 12-byte vertices: two quanta per block, runs of two
    ("color12", 2_560_000, 2560, 1, 2),  # 1000 streams x 10 blocks: the last run of a stream has two blocks, rounds straddle streams
    ("oct8", 1_000_000, 700, 1, 2),      # 1429 streams of three blocks (no full run at all)
    ("oct8", 1 << 22, 1 << 12, 0, 0),    # 1024 streams x 16 blocks, codec v0
    ("exp16", 1 << 21, 1 << 11, 1, 3),   # 1024 streams x 8 blocks, level 3 (xor / 16-bit channels)
    ("color8", 300_000, 257, 1, 2),      # streams of two blocks, the second nearly empty
])
def test_block_mode_rounds_run_major(mb, rounds_ctx, kind, count, seg, version, level):
    w = workloads.c3(kind, count=count, seg=seg, version=version, level=level)
    want = _expected(w)
    outs, status, plan, guard = device_run(w, ctx=rounds_ctx, runs=0, sidecars=_sidecars(w), block_runs=3)
    assert (status == 0).all() and guard
    assert _same(w, outs, want) is None, _same(w, outs, want)


def test_block_mode_rounds_mixed_sizes_and_bad_blocks(mb, rounds_ctx):
    """4- ... 32-byte vertices in one plan (runs of four, rounds of one / two / four members), one stream with a stale
    sidecar entry in the middle of a run and one with a framing error: the others decode bit-exact"""
    parts = [workloads.c3("oct8", count=300_000, seg=1500), workloads.c3("quat12", count=200_000, seg=999),
             workloads.c3("exp16", count=150_000, seg=4000), workloads.c1b(version=1, count=200_000),
             workloads.c2(total=1 << 17, seg=3000, level=2, version=1), workloads.c3("color8", count=100_000, seg=257, version=0, level=0)]
    w = workloads.merge("mixed", parts)
    want = _expected(w)
    sc = _sidecars(w)
    bad = [s.copy() for s in sc]
    i_stale = next(i for i in range(w.n) if int(w.vertex_sizes[i]) == 4 and int(w.counts[i]) > 4 * 256)
    i_magic = next(i for i in range(w.n) if int(w.vertex_sizes[i]) == 8)
    bad[i_stale][2] += 1                # a four-byte stream of six blocks: block 2 (inside the first run) starts one byte late
    blob = w.blob.copy()
    blob[int(w.offsets[i_magic])] = 0x55  # bad magic
    outs, status, plan, guard = device_run(w, ctx=rounds_ctx, runs=0, sidecars=bad, block_runs=2, blob_override=blob)
    assert guard
    assert status[i_stale] == mb.ERR_SIDECAR and status[i_magic] == -1
    ok = [i for i in range(w.n) if i not in (i_stale, i_magic)]
    assert (status[ok] == 0).all()
    for i in ok:
        if int(w.vertex_sizes[i]) == 4 and int(w.filters[i]) in (1, 4):
            d = np.abs(outs[i].astype(np.int16) - want[i].astype(np.int16))
            assert int(np.minimum(d, 256 - d).max(initial=0)) <= 1, i
        else:
            assert np.array_equal(outs[i], want[i]), i


def test_block_mode_rounds_level_major_still_available(mb, monkeypatch):
    """MOB200_RUN_MAJOR=0 keeps the level-major order (diagnostics / comparison runs)"""
    monkeypatch.setenv("MOB200_ROUNDS", "1")
    monkeypatch.setenv("MOB200_RUN_MAJOR", "0")
    ctx = mb.Context(-1)
    try:
        w = workloads.c3("quat12", count=400_000, seg=5000)
        outs, status, plan, guard = device_run(w, ctx=ctx, runs=0, sidecars=_sidecars(w), block_runs=2)
        assert (status == 0).all() and guard and _same(w, outs, _expected(w)) is None
    finally:
        ctx.close()


@pytest.mark.parametrize("rounds", ["0", "1", "2"])
def test_block_mode_differential_shapes(mb, monkeypatch, rounds):
    """every vertex size class x ragged counts x codec versions / levels in ONE block-mode plan with sidecars, with either
    decoder form forced and with the plan's own choice: partial blocks of large vertices join rounds (two quanta, more than
    four 4-byte lanes: never chained), streams of one block, empty streams"""
    if not loader.have_ref():
        pytest.skip("needs the reference encoder")
    R, P = loader.ref(), loader.port()
    rng = np.random.default_rng(23)
    blobs, offs, sizes, counts, vss, want, sidecars = [], [], [], [], [], [], []
    cursor, k = 0, 0
    for vs in (4, 8, 12, 16, 20, 24, 32, 48, 64, 128, 256):
        for count in (0, 1, 16, 17, 255, 256, 257, 1030, 4103):
            version, level = ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3))[k % 5]
            words = np.cumsum(rng.integers(-3, 4, (count, vs // 4)) << rng.integers(0, 20, (1, vs // 4)), axis=0).astype(np.uint32)
            v = words.view(np.uint8).reshape(-1)
            enc = R.encode_vertex_buffer(v.reshape(count, vs) if count else v, count, vs, level, version)
            rc, sc = P.block_offsets(count, vs, enc)
            assert rc == 0
            pad = (-enc.size) % 16
            blobs.append(np.concatenate([enc, np.zeros(pad, np.uint8)]))
            offs.append(cursor); sizes.append(enc.size); counts.append(count); vss.append(vs); want.append(v); sidecars.append(sc.copy())
            cursor += enc.size + pad
            k += 1
    n = len(offs)
    w = workloads.Workload("shapes", np.concatenate(blobs + [np.zeros(32, np.uint8)]), np.array(offs, np.uint64), np.array(sizes, np.uint64),
                           np.array(counts, np.uint64), np.array(vss, np.uint32), np.zeros(n, np.int32), source=None)
    monkeypatch.setenv("MOB200_ROUNDS", rounds)
    ctx = mb.Context(-1)
    try:
        outs, status, plan, guard = device_run(w, ctx=ctx, runs=0, sidecars=sidecars, block_runs=2)
        assert guard and (status == 0).all(), np.nonzero(status)[0][:8]
        for i in range(n):
            assert np.array_equal(outs[i], want[i]), (i, counts[i], vss[i], first_mismatch(outs[i], want[i]))
        del plan
    finally:
        ctx.close()


def test_block_mode_unit_chains_many_streams(mb):
    """plain form, at least as many streams as decode units: run-major order with runs of 16 blocks per unit, the decoder
    warps hand the running value from block to block (no look-back inside a run).  1500 streams x 11 blocks (ragged last
    block), one stream with a stale sidecar entry in the middle of its run, one with bad magic: the others bit-exact."""
    w = workloads.c2(total=1500 * 2700, seg=2700, level=2, version=1)
    assert w.n == 1500
    want = _expected(w)
    sc = _sidecars(w)
    bad = [s.copy() for s in sc]
    bad[700][5] += 2                      # block 5 of stream 700 starts two bytes late
    blob = w.blob.copy()
    blob[int(w.offsets[33])] = 0x10       # bad magic
    outs, status, plan, guard = device_run(w, runs=0, sidecars=bad, block_runs=2, blob_override=blob)
    assert guard
    assert status[700] == mb.ERR_SIDECAR and status[33] == -1
    ok = [i for i in range(w.n) if i not in (33, 700)]
    assert (status[ok] == 0).all()
    for i in ok:
        assert np.array_equal(outs[i], want[i]), (i, first_mismatch(outs[i], want[i]))
    # the serial walk of the same plan (level-major, no chains) gives the reference result for the stale one as well
    outs, status, plan, guard = device_run(w, runs=1)
    assert (status == 0).all() and _same(w, outs, want) is None


def test_block_mode_unit_chains_long_streams(mb):
    """800 streams x 40 blocks: a stream is three runs (16 + 16 + 8 blocks) on three different units, so every run start
    looks back across units while the blocks inside a run chain; 64-byte vertices (two quanta per warp and block)"""
    rng = np.random.default_rng(4)
    total, vs, seg = 800 * 40 * 128, 64, 40 * 128
    words = np.cumsum(rng.integers(-50, 50, (total, vs // 4)), axis=0).astype(np.uint32)
    v = words.view(np.uint8).reshape(-1)
    blob, segs, side = mb.encode_segments(v, total, vs, seg, 2, 1, threads=0)
    n = len(segs)
    assert n == 800
    w = workloads.Workload("chains64", np.concatenate([blob, np.zeros(32, np.uint8)]),
                           np.array([s.offset for s in segs], np.uint64), np.array([s.size for s in segs], np.uint64),
                           np.array([s.vertex_count for s in segs], np.uint64), np.full(n, vs, np.uint32), np.zeros(n, np.int32), source=v)
    sidecars = [side[s.sidecar_offset : s.sidecar_offset + s.sidecar_entries] for s in segs]
    outs, status, plan, guard = device_run(w, runs=0, sidecars=sidecars, block_runs=2)
    assert guard and (status == 0).all()
    assert np.array_equal(np.concatenate(outs), v)
