"""GPU parity tests (through the C ABI) for the rows next to the hot path: index-stream decode on the device
(meshopt_decodeIndexBuffer / meshopt_decodeIndexSequence and their batched form) and the glTF bufferView
front-end.  Checker: the CPU oracle (oracle/indexcodec_oracle.c, pinned against the reference in
tests/test_index_cpu.py) and the committed reference outputs under tests/golden/."""
import json
import os

import numpy as np
import pytest

from tests.index_cases import corruptions, index_sets

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KINDS = {"triangles": 0, "sequence": 1}


@pytest.fixture(scope="module")
def mb():
    import torch
    assert torch.cuda.is_available()
    import meshoptimizer_b200 as m
    m.lib()
    return m


def decode_one(mb, kind, count, size, data):
    f = mb.decode_index_buffer_rc if kind == "triangles" else mb.decode_index_sequence_rc
    return f(count, size, data)


def test_index_kat(mb, port):
    for k in json.load(open(os.path.join(ROOT, "tests", "golden", "index_kat.json"))):
        data = np.frombuffer(bytes.fromhex(k["input"]), np.uint8)
        for size in (4, 2):
            rc, out = decode_one(mb, k["kind"], k["count"], size, data)
            if k["rc"] == "negative":
                assert rc < 0 and rc == port.decode_index(k["kind"], k["count"], size, data)[0]
                continue
            want = np.frombuffer(bytes.fromhex(k["expected"]), np.uint32)
            assert rc == 0, (k["name"], rc)
            assert np.array_equal(out.astype(np.uint32), want if size == 4 else want & 0xFFFF), k["name"]


def encoded_cases(checker):
    """(kind, indices count, encoded stream): reference-encoded when oracle/_ref is present, else the KAT inputs"""
    from oracle import loader
    if loader.have_ref():
        R = loader.ref()
        for name, idx, vcount in index_sets():
            for version in (0, 1):
                for kind in KINDS:
                    yield name, kind, idx.size, R.encode_index(kind, idx, vcount, version)
    else:
        for k in json.load(open(os.path.join(ROOT, "tests", "golden", "index_kat.json"))):
            if k["rc"] == 0:
                yield k["name"], k["kind"], k["count"], np.frombuffer(bytes.fromhex(k["input"]), np.uint8)


def test_index_batch_matches_checker(mb, checker):
    """every case, both index sizes, valid / truncated / corrupted, in ONE batched launch"""
    items, want = [], []
    for ci, (name, kind, count, enc) in enumerate(encoded_cases(checker)):
        for size in (2, 4):
            variants = [enc] + list(corruptions(enc, seed=ci, n_random=6))
            for e in variants:
                items.append((e, count, size, KINDS[kind]))
                want.append(checker.decode_index(kind, count, size, e))
    outs, rcs = mb.decode_index_batch_host(items)
    assert len(items) > 100
    bad = 0
    for i, ((rc_w, out_w), rc, out) in enumerate(zip(want, rcs, outs)):
        assert rc == rc_w, (i, rc, rc_w)
        if rc == 0:
            assert np.array_equal(out, out_w), i
        bad += rc != 0
    assert bad > 10  # the error paths were exercised


def test_index_dropin_and_device_batch(mb, checker):
    import torch
    name, kind, count, enc = next(c for c in encoded_cases(checker) if c[2] > 0)
    rc, out = decode_one(mb, kind, count, 4, enc)
    rc_w, out_w = checker.decode_index(kind, count, 4, enc)
    assert rc == rc_w == 0 and np.array_equal(out, out_w)
    # device pointers: 64 copies of the stream at odd source offsets, outputs back to back
    dev = torch.device("cuda:0")
    n = 64
    pitch = (enc.size + 1 + 15) & ~15
    src = torch.zeros(n * pitch + 16, dtype=torch.uint8, device=dev)
    for i in range(n):
        src[i * pitch + (i & 1) : i * pitch + (i & 1) + enc.size] = torch.from_numpy(np.ascontiguousarray(enc)).to(dev)
    dst = torch.zeros(n * count, dtype=torch.int32, device=dev)
    arr = (mb.IndexStream * n)()
    for i in range(n):
        arr[i].src = src.data_ptr() + i * pitch + (i & 1)
        arr[i].src_size = enc.size
        arr[i].dst = dst.data_ptr() + 4 * i * count
        arr[i].index_count = count
        arr[i].index_size = 4
        arr[i].kind = KINDS[kind]
    rc = mb.lib().mob200_decode_index_batch_device(mb.default_context().handle, arr, n, None)
    assert rc == 0
    got = dst.cpu().numpy().view(np.uint32).reshape(n, count)
    assert (got == out_w[None, :]).all()


def test_index_argument_errors(mb):
    with pytest.raises(ValueError):
        mb.decode_index_buffer_rc(4, 4, b"\xe1" + bytes(32))
    arr = (mb.IndexStream * 1)()
    arr[0].src = None
    arr[0].index_count = 3
    arr[0].index_size = 3
    rc = mb.lib().mob200_decode_index_batch_host(mb.default_context().handle, arr, 1)
    assert rc == 1 and arr[0].status == mb.ERR_ARGUMENT
    rc, _ = mb.decode_index_sequence_rc(5, 2, b"")
    assert rc == -2


@pytest.mark.parametrize("name", ["c", "cc", "cc_float", "khr"])
def test_gltf_fixture_decodes_like_reference(mb, name):
    blob = np.fromfile(os.path.join(ROOT, "tests", "golden", f"gltf_{name}.glb"), dtype=np.uint8)
    expected = np.load(os.path.join(ROOT, "tests", "golden", "gltf_expected.npz"))
    outputs, views, info = mb.gltf_decode_host(blob)
    assert info.view_count >= 3 and all(v.status == 0 for v in list(views)[: info.view_count])
    assert {v.mode for v in list(views)[: info.view_count]} >= {mb.GLTF_ATTRIBUTES, mb.GLTF_TRIANGLES}
    for b, got in outputs.items():
        want = expected[f"{name}_buffer{b}"]
        # octahedral normals stored in 4 bytes take the <=1 LSB lane; everything else is bit-exact
        exact = np.ones(want.size, bool)
        for v in list(views)[: info.view_count]:
            if v.dst_buffer == b and v.filter == mb.FILTER_OCTAHEDRAL and v.stride == 4:
                exact[v.dst_offset : v.dst_offset + v.dst_size] = False
        assert np.array_equal(got[exact], want[exact]), name
        d = np.abs(got[~exact].astype(np.int16) - want[~exact].astype(np.int16))
        assert d.size == 0 or int(np.minimum(d, 256 - d).max()) <= 1


def test_gltf_device_variant_and_bad_view(mb):
    import ctypes

    import torch
    blob = np.fromfile(os.path.join(ROOT, "tests", "golden", "gltf_c.glb"), dtype=np.uint8)
    expected = np.load(os.path.join(ROOT, "tests", "golden", "gltf_expected.npz"))["c_buffer1"]
    views, sizes, info = mb.gltf_scan(blob)
    dev = torch.device("cuda:0")
    d_bin = torch.from_numpy(blob[info.bin_offset : info.bin_offset + info.bin_size].copy()).to(dev)
    d_out = torch.zeros(sizes[1] + 16, dtype=torch.uint8, device=dev)
    bufs = (ctypes.c_void_p * 2)(d_bin.data_ptr(), None)
    outs = (ctypes.c_void_p * 2)(None, d_out.data_ptr())
    lens = (ctypes.c_size_t * 2)(info.bin_size, 0)
    olens = (ctypes.c_size_t * 2)(0, sizes[1])
    rc = mb.lib().mob200_gltf_decode_device(mb.default_context().handle, views, info.view_count, 2, bufs, lens, outs, olens, None)
    assert rc == 0
    assert np.array_equal(d_out.cpu().numpy()[: sizes[1]], expected)
    # a view whose compressed range is cut short fails with the codec's own code; the others still decode
    views2, _, _ = mb.gltf_scan(blob)
    views2[0].src_size -= 5
    rc = mb.lib().mob200_gltf_decode_device(mb.default_context().handle, views2, info.view_count, 2, bufs, lens, outs, olens, None)
    assert rc == 1 and views2[0].status in (-2, -3) and all(views2[i].status == 0 for i in range(1, info.view_count))
    # views are untrusted: a buffer index beyond the asset's buffers, a destination range beyond its buffer and a
    # source range beyond its buffer are rejected before anything is read or written (guard bytes stay)
    views3, _, _ = mb.gltf_scan(blob)
    views3[0].dst_buffer = 7
    views3[1].dst_offset = sizes[1] - 4
    views3[2].src_offset = info.bin_size - 3
    d_out.fill_(0xCD)
    rc = mb.lib().mob200_gltf_decode_device(mb.default_context().handle, views3, info.view_count, 2, bufs, lens, outs, olens, None)
    assert rc == 3 and all(views3[i].status == mb.ERR_ARGUMENT for i in range(3)) and all(views3[i].status == 0 for i in range(3, info.view_count))
    assert (d_out.cpu().numpy()[sizes[1]:] == 0xCD).all()


# ---- meshlets ---------------------------------------------------------------------------------------

def test_meshlet_batch_matches_checker(mb, checker):
    """encoder-produced meshlets in all four output formats, plus truncated / corrupted ones, in ONE launch"""
    from oracle import loader
    from tests.meshlet_cases import corruptions as meshlet_corruptions, meshlets
    if not loader.have_ref():
        pytest.skip("the reference encoder (oracle/_ref) is needed to produce meshlets")
    R = loader.ref()
    items, want = [], []
    for ci, (name, verts, tris) in enumerate(meshlets()):
        enc = R.encode_meshlet(verts, tris)
        for vs in (2, 4):
            for ts in (3, 4):
                for e in [enc] + list(meshlet_corruptions(enc, seed=ci, n_random=6)):
                    items.append((e, verts.size, vs, tris.shape[0], ts))
                    want.append(checker.decode_meshlet(verts.size, vs, tris.shape[0], ts, e))
    outs, rcs = mb.decode_meshlet_batch_host(items)
    bad = 0
    for i, ((rc_w, v_w, t_w), rc, (v, t)) in enumerate(zip(want, rcs, outs)):
        assert rc == rc_w, (i, rc, rc_w)
        if rc == 0:
            assert np.array_equal(v, v_w) and np.array_equal(t, t_w), i
        bad += rc != 0
    assert len(items) > 300 and bad > 20


def test_meshlet_dropin(mb, checker):
    from oracle import loader
    from tests.meshlet_cases import meshlets
    if not loader.have_ref():
        pytest.skip("the reference encoder (oracle/_ref) is needed to produce meshlets")
    name, verts, tris = next(iter(meshlets()))
    enc = loader.ref().encode_meshlet(verts, tris)
    rc, v, t = mb.decode_meshlet_rc(verts.size, 4, tris.shape[0], 3, enc)
    rc_w, v_w, t_w = checker.decode_meshlet(verts.size, 4, tris.shape[0], 3, enc)
    assert rc == rc_w == 0 and np.array_equal(v, v_w) and np.array_equal(t, t_w)
    rc, _, _ = mb.decode_meshlet_rc(verts.size, 2, tris.shape[0], 4, enc[:10])
    assert rc == checker.decode_meshlet(verts.size, 2, tris.shape[0], 4, enc[:10])[0] == -2
    with pytest.raises(ValueError):
        mb.decode_meshlet_rc(300, 4, 10, 3, enc)
