"""CPU tests of the index-stream checker (oracle/indexcodec_oracle.c): the reference's known-answer vectors,
and -- when oracle/_ref is present -- a differential run against the unmodified reference decoders on
encoder-produced, truncated and corrupted streams.  Also the host logic of the glTF scanner (no GPU)."""
import json
import os

import numpy as np
import pytest

from tests.index_cases import corruptions, index_sets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def index_kat():
    with open(os.path.join(ROOT, "tests", "golden", "index_kat.json")) as f:
        return json.load(f)


def check_kat(lib, kat):
    for k in kat:
        data = np.frombuffer(bytes.fromhex(k["input"]), np.uint8)
        for size in (4, 2):
            rc, out = lib.decode_index(k["kind"], k["count"], size, data)
            if k["rc"] == "negative":
                assert rc < 0, (k["name"], rc)
                continue
            want = np.frombuffer(bytes.fromhex(k["expected"]), np.uint32)
            assert rc == 0, (k["name"], rc)
            assert np.array_equal(out.astype(np.uint32), want if size == 4 else want & 0xFFFF), k["name"]


def test_port_reproduces_reference_kat(port, index_kat):
    check_kat(port, index_kat)


def test_reference_reproduces_its_own_kat(index_kat):
    from oracle import loader
    if not loader.have_ref():
        pytest.skip("oracle/_ref not built")
    check_kat(loader.ref(), index_kat)


def test_version_probe(port):
    assert port.decode_index_version(bytes([0xE1, 0])) == 1
    assert port.decode_index_version(bytes([0xD0])) == 0
    assert port.decode_index_version(bytes([0xE2])) == -1
    assert port.decode_index_version(bytes([0xA1])) == -1
    assert port.decode_index_version(b"") == -1
    import meshoptimizer_b200 as mb  # host logic of the product library: identical answers, no device needed
    for probe in (bytes([0xE1, 0]), bytes([0xD0]), bytes([0xE2]), bytes([0xA1]), b""):
        assert mb.decode_index_version(probe) == port.decode_index_version(probe)


def test_port_matches_reference_differential(port):
    from oracle import loader
    if not loader.have_ref():
        pytest.skip("oracle/_ref not built")
    R = loader.ref()
    cases = 0
    for name, idx, vcount in index_sets():
        for version in (0, 1):
            for kind in ("triangles", "sequence"):
                enc = R.encode_index(kind, idx, vcount, version)
                for size in (2, 4):
                    a, b = R.decode_index(kind, idx.size, size, enc), port.decode_index(kind, idx.size, size, enc)
                    assert a[0] == b[0] == 0 and np.array_equal(a[1], b[1]), (name, version, kind, size)
                    want = idx if size == 4 else idx & 0xFFFF
                    if kind == "sequence":
                        assert np.array_equal(a[1].astype(np.uint32), want)
                    else:  # the triangle codec may rotate a triangle (same winding, different first corner)
                        t, w = a[1].astype(np.uint32).reshape(-1, 3), want.reshape(-1, 3)
                        assert ((t == w).all(1) | (t == np.roll(w, 1, 1)).all(1) | (t == np.roll(w, 2, 1)).all(1)).all()
                    for e in corruptions(enc, seed=cases):
                        a, b = R.decode_index(kind, idx.size, size, e), port.decode_index(kind, idx.size, size, e)
                        assert a[0] == b[0], (name, version, kind, size, a[0], b[0])
                        if a[0] == 0:
                            assert np.array_equal(a[1], b[1])
                        cases += 1
    assert cases > 1000


def test_gltf_scan_fixtures():
    """host logic of the bufferView front-end: the scanner finds what Python's json module finds"""
    import struct

    import meshoptimizer_b200 as mb

    for name in ("c", "cc", "cc_float", "khr"):
        blob = open(os.path.join(ROOT, "tests", "golden", f"gltf_{name}.glb"), "rb").read()
        views, sizes, info = mb.gltf_scan(blob)
        n, t = struct.unpack_from("<II", blob, 12)
        doc = json.loads(blob[20 : 20 + n])
        assert (info.json_offset, info.json_size) == (20, n)
        assert sizes == [b["byteLength"] for b in doc["buffers"]]
        want = []
        for i, v in enumerate(doc["bufferViews"]):
            ext = v.get("extensions", {})
            mc = ext.get("EXT_meshopt_compression") or ext.get("KHR_meshopt_compression")
            if mc:
                want.append((i, mc["mode"], mc.get("filter", "NONE"), mc["buffer"], mc.get("byteOffset", 0), mc["byteLength"], mc["count"], mc["byteStride"],
                             v["buffer"], v.get("byteOffset", 0), v["byteLength"]))
        assert info.view_count == len(want) and info.invalid_views == 0
        modes = ["ATTRIBUTES", "TRIANGLES", "INDICES"]
        filters = ["NONE", "OCTAHEDRAL", "QUATERNION", "EXPONENTIAL", "COLOR"]
        got = [(v.view, modes[v.mode], filters[v.filter], v.src_buffer, v.src_offset, v.src_size, v.count, v.stride, v.dst_buffer, v.dst_offset, v.dst_size)
               for v in list(views)[: info.view_count]]
        assert got == want
        assert info.bin_size == sizes[0] or info.bin_size == ((sizes[0] + 3) & ~3)


def test_gltf_scan_rejects_garbage():
    import meshoptimizer_b200 as mb

    for bad in (b"glTF\x02\x00\x00\x00\xff\xff\xff\x7f", b"{\"bufferViews\":[{\"buffer\":", b"[1,2", b"glTF\x01\x00\x00\x00\x0c\x00\x00\x00"):
        with pytest.raises(ValueError):
            mb.gltf_scan(bad)
    # rule violations are reported per view, not as a parse error (extern/cgltf.h:1645-1667)
    doc = {"buffers": [{"byteLength": 64}, {"byteLength": 96}],
           "bufferViews": [{"buffer": 1, "byteLength": 96, "extensions": {"EXT_meshopt_compression": {"buffer": 0, "byteLength": 64, "byteStride": 6, "count": 16, "mode": "ATTRIBUTES"}}},
                           {"buffer": 1, "byteLength": 32, "extensions": {"EXT_meshopt_compression": {"buffer": 0, "byteLength": 64, "byteStride": 4, "count": 8, "mode": "TRIANGLES"}}},
                           {"buffer": 1, "byteLength": 64, "extensions": {"KHR_meshopt_compression": {"buffer": 0, "byteLength": 64, "byteStride": 4, "count": 16, "mode": "ATTRIBUTES", "filter": "QUATERNION"}}},
                           {"buffer": 1, "byteLength": 64, "extensions": {"KHR_meshopt_compression": {"buffer": 0, "byteLength": 64, "byteStride": 4, "count": 16, "mode": "ATTRIBUTES", "filter": "OCTAHEDRAL"}}}]}
    views, sizes, info = mb.gltf_scan(json.dumps(doc).encode())
    assert info.view_count == 4 and info.invalid_views == 3
    assert [v.status for v in list(views)[:4]] == [mb.ERR_ARGUMENT, mb.ERR_ARGUMENT, mb.ERR_ARGUMENT, 0]
