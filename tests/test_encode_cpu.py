"""The segmenter / encoder-side helper (include/meshopt_b200.h section 2c), host code.

* every stream it makes is accepted by the UNMODIFIED reference decoder and decodes to the original vertices
  (reference src/vertexcodec.cpp:1799-1872), and by the pinned port when the reference library is absent;
* it is byte-identical to the reference encoder at the same version / level (:1615-1693), so ratios are the same;
* the sidecars it emits are the block offsets the CPU walk finds;
* mob200_encode_segments: K independent streams, 16-byte packed, descriptors and sidecars consistent.
"""
import numpy as np
import pytest

from oracle import loader

SIZES = (4, 8, 12, 16, 20, 32, 48, 64, 128, 256)
COUNTS = (0, 1, 13, 16, 17, 255, 256, 257, 1000, 4103)
MODES = ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3))


@pytest.fixture(scope="module")
def mb():
    import meshoptimizer_b200 as m

    m.lib()
    return m


def vertex_data(kind, count, vs, rng):
    if kind == 0:
        return rng.integers(0, 256, (count, vs), dtype=np.uint8)
    if kind == 1:
        return np.cumsum(rng.integers(-3, 4, (count, vs)), axis=0).astype(np.uint8)
    if kind == 2:
        return np.cumsum(rng.integers(-300, 300, (count, vs // 4)), axis=0).astype(np.uint32).view(np.uint8).reshape(count, vs)
    if kind == 3:
        return (0xF0 + np.cumsum(rng.integers(-5, 9, (count, vs // 2)), axis=0)).astype(np.uint16).view(np.uint8).reshape(count, vs)
    i = np.arange(count, dtype=np.uint64)[:, None]
    sh = (np.arange(vs // 4, dtype=np.uint64) * 7 % 29)[None, :]
    return ((i * 3 + rng.integers(0, 3, (count, vs // 4)).astype(np.uint64)) << sh).astype(np.uint32).view(np.uint8).reshape(count, vs)


def test_streams_decode_with_the_checker_and_match_the_reference_encoder(mb, checker, port):
    rng = np.random.default_rng(21)
    have_ref = loader.have_ref()
    R = loader.ref() if have_ref else None
    n = 0
    for vs in SIZES:
        for count in COUNTS:
            for version, level in MODES:
                v = vertex_data(n % 5, count, vs, rng)
                n += 1
                enc, side = mb.encode_vertex_buffer(v, count, vs, level, version, with_sidecar=True)
                assert enc.size <= mb.lib().mob200_encode_vertex_bound(count, vs)
                rc, dec = checker.decode_vertex_buffer(count, vs, enc)
                assert rc == 0 and np.array_equal(dec, v.reshape(-1)), (vs, count, version, level)
                rc, off = port.block_offsets(count, vs, enc)
                assert rc == 0 and np.array_equal(off, side), (vs, count, version, level)
                if have_ref:
                    assert np.array_equal(enc, R.encode_vertex_buffer(v, count, vs, level, version)), (vs, count, version, level)


def test_encode_reports_short_buffers_and_bad_arguments(mb):
    v = np.arange(64 * 16, dtype=np.uint8)
    L = mb.lib()
    buf = np.zeros(4096, np.uint8)
    full = L.mob200_encode_vertex_buffer(buf.ctypes.data, buf.size, v.ctypes.data, 64, 16, 2, 1, None)
    assert full > 0
    for cap in (0, 1, 10, full - 1):
        assert L.mob200_encode_vertex_buffer(buf.ctypes.data, cap, v.ctypes.data, 64, 16, 2, 1, None) == 0
    assert L.mob200_encode_vertex_buffer(buf.ctypes.data, buf.size, v.ctypes.data, 64, 6, 2, 1, None) == 0   # vertex size not a multiple of 4
    assert L.mob200_encode_vertex_buffer(buf.ctypes.data, buf.size, v.ctypes.data, 64, 16, 2, 2, None) == 0  # no such version
    assert L.mob200_encode_vertex_bound(10, 6) == 0


@pytest.mark.parametrize("vs,total,seg,version,level", [(32, 100_003, 4096, 1, 2), (16, 70_000, 65_536, 0, 0), (12, 5_000, 0, 1, 3), (4, 33_333, 1000, 1, 2), (64, 9_999, 512, 1, 1)])
def test_segments(mb, checker, port, vs, total, seg, version, level):
    rng = np.random.default_rng(5)
    v = vertex_data(2 if vs >= 8 else 1, total, vs, rng).reshape(-1)
    blob, segs, side = mb.encode_segments(v, total, vs, seg, level, version, threads=4)
    assert len(segs) == (1 if seg == 0 else (total + seg - 1) // seg)
    at = 0
    cursor = 0
    for i, s in enumerate(segs):
        assert s.first_vertex == at and s.offset % 16 == 0 and s.offset >= cursor
        cursor = s.offset + s.size
        stream = blob[s.offset : s.offset + s.size]
        rc, dec = checker.decode_vertex_buffer(s.vertex_count, vs, stream)
        assert rc == 0 and np.array_equal(dec, v[at * vs : (at + s.vertex_count) * vs]), i
        # an independent stream of its own: identical to encoding that range alone
        assert np.array_equal(stream, mb.encode_vertex_buffer(v[at * vs : (at + s.vertex_count) * vs], s.vertex_count, vs, level, version))
        rc, off = port.block_offsets(s.vertex_count, vs, stream)
        assert s.sidecar_entries == off.size and np.array_equal(side[s.sidecar_offset : s.sidecar_offset + s.sidecar_entries], off)
        at += s.vertex_count
    assert at == total and cursor <= blob.size
