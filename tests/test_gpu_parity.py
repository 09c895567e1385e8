"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU checker.

Bar: every codec byte identical; Exp / Oct16 / Quat / Color16 identical;
Oct8 and Color8 within 1 LSB (the reference uses rsqrtps / rcpps there) with the passthrough byte exact.
Nothing here reads /root/reference: streams come from the reference encoder inside oracle/_ref (a prebuilt
library that travels with the repository) or from committed fixtures.
"""
import numpy as np
import pytest

from oracle import loader, workloads
from tests.gpu_util import device_run, first_mismatch

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not present (reference encoder needed to generate streams)")


@pytest.fixture(scope="module")
def mb():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import meshoptimizer_b200 as m
    m.lib()  # fails loudly if the extension is missing
    return m


def _close_8bit(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int(np.minimum(d, 256 - d).max()) if d.size else 0


def _check_filtered(kind_or_name, stride, got, want, ctx=""):
    if stride == 4 and kind_or_name in ("oct", "color", "oct8", "color8"):
        assert _close_8bit(got, want) <= 1, (ctx, "8-bit tolerance lane off by more than 1")
        if kind_or_name in ("oct", "oct8"):
            assert np.array_equal(got[3::4], want[3::4]), (ctx, "W passthrough")
    else:
        assert np.array_equal(got, want), (ctx, first_mismatch(got, want))


# ---- reference known-answer vectors through the drop-in host API -------------------------------------------

def test_kat_codec_host(mb, kat):
    for k in kat["codec"]:
        rc, out = mb.decode_vertex_buffer_rc(k["count"], k["size"], bytes.fromhex(k["input"]))
        assert rc == k["rc"], k["name"]
        assert out.tobytes().hex() == k["expected"], k["name"]


def test_kat_filters_host(mb, kat):
    fn = {"oct": mb.decode_filter_oct, "quat": mb.decode_filter_quat, "exp": mb.decode_filter_exp, "color": mb.decode_filter_color}
    for k in kat["filter"]:
        for count in (4, 3):
            buf = np.frombuffer(bytes.fromhex(k["input"]), dtype=np.uint8)[: count * k["stride"]].copy()
            fn[k["filter"]](buf, count, k["stride"])
            assert buf.tobytes().hex() == k["expected"][: count * k["stride"] * 2], (k["name"], count)


def test_kat_fused(mb, kat):
    for k in kat["fused"]:
        out = mb.decode_vertex_buffer(k["count"], k["size"], bytes.fromhex(k["input"]), filter=k["filter"])
        assert out.tobytes().hex() == k["expected"], k["name"]


def test_kat_version(mb, kat):
    for k in kat["version"]:
        assert mb.decode_vertex_version(bytes.fromhex(k["input"])) == k["rc"]


# ---- committed reference outputs ---------------------------------------------------------------------------------

def test_reference_fixtures_device(mb, ref_vectors):
    meta = ref_vectors["codec_meta"]
    items = [(ref_vectors[f"codec_{i}_enc"], int(c), int(vs), 0) for i, (c, vs, _, _) in enumerate(meta)]
    outs, rcs = mb.decode_batch_host(items)
    for i, (c, vs, version, level) in enumerate(meta):
        want = ref_vectors[f"codec_{i}_dec"]
        assert rcs[i] == 0, (i, c, vs, version, level)
        assert np.array_equal(outs[i], want), (i, int(c), int(vs), int(version), int(level), first_mismatch(outs[i], want))


def test_reference_fixtures_filters(mb, ref_vectors):
    import torch
    for kind, fname, stride, count in ref_vectors["filter_meta"]:
        stride, count, fname = int(stride), int(count), str(fname)
        want = ref_vectors[f"filter_{kind}_out"]
        # standalone device filter
        d = torch.from_numpy(ref_vectors[f"filter_{kind}_in"].copy()).cuda()
        mb.filter_device(fname, d.data_ptr(), count, stride, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        _check_filtered(str(kind), stride, d.cpu().numpy(), want, kind)
        # host drop-in with a count that is not a multiple of 4
        h = ref_vectors[f"filter_{kind}_in"][: (count - 1) * stride].copy()
        getattr(mb, "decode_filter_" + fname)(h, count - 1, stride)
        _check_filtered(str(kind), stride, h, want[: (count - 1) * stride], kind)


# ---- differential against the CPU checker on encoder-produced streams ----------------------------------------------

def _vertex_data(kind, count, vs, rng):
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


@needs_ref
def test_differential_batch(mb, checker):
    """vertex sizes x counts x versions x levels in ONE batched launch (ragged, incl. empty streams)"""
    R = loader.ref()
    rng = np.random.default_rng(11)
    items, want = [], []
    n = 0
    for vs in (4, 8, 12, 16, 20, 24, 32, 48, 64, 128, 256):
        for count in (0, 1, 13, 16, 17, 255, 256, 257, 4103):
            for version, level in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3)):
                v = _vertex_data(n % 5, count, vs, rng)
                enc = R.encode_vertex_buffer(v, count, vs, level, version)
                items.append((enc, count, vs, 0))
                want.append(v.reshape(-1))
                n += 1
    outs, rcs = mb.decode_batch_host(items)
    for i, (it, w) in enumerate(zip(items, want)):
        assert rcs[i] == 0, (i, it[1], it[2])
        assert np.array_equal(outs[i], w), (i, it[1], it[2], first_mismatch(outs[i], w))


@needs_ref
@pytest.mark.parametrize("dst_shift,src_shift", [(0, 0), (4, 1), (1, 7), (8, 3)])
def test_alignment_paths(mb, dst_shift, src_shift):
    """arbitrary source alignment (glTF aligns views to 4) and the 16/4/1-byte store paths"""
    w = workloads.c2(total=20_000, seg=3_001, level=3, version=1)
    outs, status, _, guard = device_run(w, dst_shift=dst_shift, src_shift=src_shift)
    assert (status == 0).all()
    assert guard, "bytes outside the destination ranges were written"
    got = np.concatenate(outs)
    assert np.array_equal(got, w.source), first_mismatch(got, w.source)


@needs_ref
def test_error_codes_and_memory_safety(mb, checker):
    """truncations, trailing bytes, broken headers, invalid channel byte: same return code as the
    reference for every case, and the batch never hangs (demo/tests.cpp:521-572)"""
    R = loader.ref()
    rng = np.random.default_rng(3)
    items, want_rc = [], []
    for vs, count, version in ((12, 4, 1), (12, 4, 0), (4, 13, 1), (32, 300, 1), (16, 257, 0)):
        v = _vertex_data(1, count, vs, rng)
        enc = R.encode_vertex_buffer(v, count, vs, 2, version)
        cuts = list(range(0, min(enc.size, 70))) + list(range(max(0, enc.size - 70), enc.size + 1))
        for cut in cuts:
            items.append((enc[:cut].copy(), count, vs, 0))
        items.append((np.concatenate([enc, np.zeros(1, np.uint8)]), count, vs, 0))
        broken = enc.copy(); broken[0] = 0
        items.append((broken, count, vs, 0))
        if version == 1:
            bad = enc.copy(); bad[-1] |= 3  # channel mode 3 is invalid (vertexcodec.cpp:1584-1585)
            items.append((bad, count, vs, 0))
    for it in range(400):  # garbage with a valid magic (tools/codecfuzz.cpp shape)
        vs = [4, 16, 24, 32][it % 4]
        ln = int(rng.integers(1, 400))
        g = rng.integers(0, 256, ln, dtype=np.uint8)
        g[0] = 0xA0 | (it & 1)
        if it % 3 == 0 and ln > 4:
            g[1 : ln // 2] = rng.integers(0, 4, ln // 2 - 1, dtype=np.uint8)
        items.append((g, 66, vs, 0))
    for it in items:
        want_rc.append(checker.decode_vertex_buffer(it[1], it[2], it[0]))
    outs, rcs = mb.decode_batch_host(items)
    for i, (it, (rc, out)) in enumerate(zip(items, want_rc)):
        assert rcs[i] == rc, (i, it[1], it[2], len(it[0]), rcs[i], rc)
        if rc == 0:
            assert np.array_equal(outs[i], out), (i, first_mismatch(outs[i], out))


# ---- the BASELINE.json configs at test sizes ------------------------------------------------------------------------

@needs_ref
@pytest.mark.parametrize("version,level", [(0, 0), (1, 2)])
def test_c1a_grid(mb, version, level):
    w = workloads.c1a(version=version, level=level, side=300)
    outs, status, _, guard = device_run(w)
    assert (status == 0).all() and guard
    assert np.array_equal(outs[0], w.source), first_mismatch(outs[0], w.source)


@needs_ref
@pytest.mark.parametrize("version", [0, 1])
def test_c1b_js16(mb, version):
    w = workloads.c1b(version=version, count=1 << 17)
    outs, status, _, guard = device_run(w)
    assert (status == 0).all() and guard
    assert np.array_equal(outs[0], w.source), first_mismatch(outs[0], w.source)


@needs_ref
@pytest.mark.parametrize("level,version,seg", [(2, 1, 1 << 12), (3, 1, 1 << 14), (0, 0, 5000), (2, 1, None)])
def test_c2_segments(mb, level, version, seg):
    w = workloads.c2(total=1 << 18, seg=seg, level=level, version=version)
    outs, status, _, guard = device_run(w, runs=2)  # second run exercises the look-back epoch
    assert (status == 0).all() and guard
    got = np.concatenate(outs)
    assert np.array_equal(got, w.source), first_mismatch(got, w.source)


@needs_ref
@pytest.mark.parametrize("kind", workloads.C3_KINDS)
@pytest.mark.parametrize("version", [0, 1])
def test_c3_fused_filters(mb, checker, kind, version):
    w = workloads.c3(kind, count=70_001, seg=9_000, version=version, level=2 if version else 0)
    want = workloads.expected_outputs(w, lib=checker)
    outs, status, _, guard = device_run(w)
    assert (status == 0).all() and guard
    stride = int(w.vertex_sizes[0])
    for i in range(w.n):
        _check_filtered(kind, stride, outs[i], want[i], (kind, version, i))


@needs_ref
def test_c4_small_streams(mb, checker):
    w = workloads.c4(5_000)
    want = workloads.expected_outputs(w, lib=checker)
    outs, status, _, guard = device_run(w)
    assert (status == 0).all() and guard
    for i in range(w.n):
        assert np.array_equal(outs[i], want[i]), (i, int(w.counts[i]), int(w.vertex_sizes[i]), first_mismatch(outs[i], want[i]))


# ---- small-vertex batches: the rounds form of the decode roles (up to four blocks per unit at a time) --------------------

def _run_and_check(mb, w, checker, ctx=None, tag="", runs=2):
    want = workloads.expected_outputs(w, lib=checker)
    outs, status, _, guard = device_run(w, ctx=ctx, runs=runs)
    assert (status == 0).all() and guard, tag
    names = {v: k for k, v in loader.FILTER_NAMES.items()}
    for i in range(w.n):
        _check_filtered(names.get(int(w.filters[i]), "none"), int(w.vertex_sizes[i]), outs[i], want[i], (tag, i))


@pytest.fixture
def rounds_ctx(mb, monkeypatch):
    """a context with the rounds form forced on (the plan heuristic only picks it for many-stream batches)"""
    monkeypatch.setenv("MOB200_ROUNDS", "1")
    ctx = mb.Context(-1)
    yield ctx
    ctx.close()


@needs_ref
@pytest.mark.parametrize("kind,count,seg", [
    ("oct8", 800_000, None),      # one stream of 3125 four-byte blocks: every unit decodes rounds of four blocks of the SAME stream
    ("quat12", 800_000, None),    # the same with 8-byte vertices
    ("exp15", 900_000, None),     # 12-byte vertices: two quanta per block, rounds of two
    ("color12", 2_560_000, 2560), # 1000 streams x 10 blocks: a block's predecessor sits one or two places earlier in another unit's queue
    ("oct8", 1_000_000, 700),     # 1429 streams of three blocks (last one ragged)
    ("oct8", 1 << 22, 1 << 12),   # 1024 streams x 16 blocks
    ("exp16", 1 << 21, 1 << 11),  # 1024 streams x 8 blocks, 12-byte vertices
])
def test_rounds_small_vertices(mb, checker, rounds_ctx, kind, count, seg):
    w = workloads.c3(kind, count=count, seg=seg, version=1, level=2)
    _run_and_check(mb, w, checker, ctx=rounds_ctx, tag=(kind, count, seg), runs=6)


@needs_ref
@pytest.mark.parametrize("kind,count,seg", [("oct8", 1 << 23, 256), ("quat12", 1 << 23, 512), ("exp15", 1 << 22, 300), ("color12", 1 << 22, 1024)])
def test_rounds_many_streams(mb, checker, kind, count, seg):
    """the regime the plan heuristic picks the rounds form for (decoders are the bottleneck, the producer runs ahead):
    every byte compared"""
    w = workloads.c3(kind, count=count, seg=seg, version=1, level=2)
    _run_and_check(mb, w, checker, tag=(kind, count, seg), runs=6)


@needs_ref
@pytest.mark.parametrize("force", ["0", "1"])
def test_rounds_mixed_vertex_sizes(mb, checker, force, monkeypatch):
    """4-, 8-, 12-, 16- and 32-byte vertices in one batch, filters on some, with either decoder form forced."""
    parts = [workloads.c3("oct8", count=300_000, seg=1500), workloads.c3("quat12", count=200_000, seg=999),
             workloads.c3("exp16", count=150_000, seg=4000), workloads.c1b(version=1, count=200_000),
             workloads.c2(total=1 << 17, seg=3000, level=2, version=1), workloads.c3("color8", count=100_000, seg=257, version=0, level=0)]
    w = workloads.merge("mixed", parts)
    monkeypatch.setenv("MOB200_ROUNDS", force)
    ctx = mb.Context(-1)
    try:
        _run_and_check(mb, w, checker, ctx=ctx, tag=("mixed", force))
    finally:
        ctx.close()


@needs_ref
def test_full_size_roundtrip_property(mb):
    """size-independent property at a large size: decode(encode(x)) == x for 16 Mi vertices (512 MB),
    compared on the device (checksum-free exact compare)"""
    import torch
    w = workloads.c2(total=1 << 24, seg=1 << 13, level=2, version=1)
    outs, status, plan, guard = device_run(w)
    assert (status == 0).all() and guard
    got = np.concatenate(outs)
    assert got.size == w.source.size
    assert np.array_equal(got, w.source)


def test_host_api_roundtrip_without_ref(mb, ref_vectors):
    """drop-in host call on a committed fixture (works even when oracle/_ref is absent)"""
    i = 40
    c, vs, _, _ = ref_vectors["codec_meta"][i]
    out = mb.decode_vertex_buffer(int(c), int(vs), ref_vectors[f"codec_{i}_enc"])
    assert np.array_equal(out, ref_vectors[f"codec_{i}_dec"])


@needs_ref
@pytest.mark.parametrize("pinned", [False, True])
def test_host_batch_chunked_pipeline(mb, pinned):
    """mob200_decode_batch_host on ~400 MB of traffic: several chunks in flight over the three slots,
    merged host->device / device->host copies, pageable (staged) and pinned (in place) caller memory"""
    import torch
    w = workloads.c2(total=1 << 23, seg=1 << 12, level=2, version=1)
    out_offs = w.out_offsets()
    if pinned:
        h_in = torch.from_numpy(w.blob).pin_memory()
        h_out = torch.zeros(w.out_bytes() + 64, dtype=torch.uint8).pin_memory()
        in_ptr, out_ptr, out_np = h_in.data_ptr(), h_out.data_ptr(), h_out.numpy()
    else:
        h_out_np = np.zeros(w.out_bytes() + 64, dtype=np.uint8)
        in_ptr, out_ptr, out_np = w.blob.ctypes.data, h_out_np.ctypes.data, h_out_np
    items = [(in_ptr + int(w.offsets[i]), int(w.sizes[i]), out_ptr + int(out_offs[i]), int(w.counts[i]), 32, 0) for i in range(w.n)]
    arr = mb.make_streams(items)
    rc = mb.lib().mob200_decode_batch_host(mb.default_context().handle, arr, w.n)
    assert rc == 0
    assert all(arr[i].status == 0 for i in range(w.n))
    got = np.concatenate([out_np[int(out_offs[i]) : int(out_offs[i]) + int(w.counts[i]) * 32] for i in range(w.n)])
    assert np.array_equal(got, w.source), first_mismatch(got, w.source)
