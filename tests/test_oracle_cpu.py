"""CPU tests: pin the oracle (C restatement) to the reference's own vectors and outputs.

Mirrors the reference test tiers: known-answer vectors from demo/tests.cpp and
js/meshopt_decoder.test.js, return-code / memory-safety tests (demo/tests.cpp:521-572,743-751),
reference-generated fixtures, and -- when oracle/_ref is present -- a differential run against the
unmodified reference decoder.
"""
import numpy as np
import pytest

from oracle import loader


def _libs():
    loader.build()
    libs = [loader.port()]
    if loader.have_ref():
        libs.append(loader.ref())
    return libs


@pytest.mark.parametrize("lib", _libs(), ids=lambda l: l.kind)
def test_codec_kat(lib, kat):
    for k in kat["codec"]:
        rc, out = lib.decode_vertex_buffer(k["count"], k["size"], bytes.fromhex(k["input"]))
        assert rc == k["rc"], k["name"]
        assert out.tobytes().hex() == k["expected"], k["name"]


@pytest.mark.parametrize("lib", _libs(), ids=lambda l: l.kind)
def test_filter_kat(lib, kat):
    for k in kat["filter"]:
        for count in (4, 3):  # 3 hits the SIMD tail path of the reference (vertexfilter.cpp:217-241)
            buf = np.frombuffer(bytes.fromhex(k["input"]), dtype=np.uint8)[: count * k["stride"]]
            out = lib.decode_filter(k["filter"], buf, count, k["stride"])
            assert out.tobytes().hex() == k["expected"][: count * k["stride"] * 2], (k["name"], count)


@pytest.mark.parametrize("lib", _libs(), ids=lambda l: l.kind)
def test_fused_kat(lib, kat):
    for k in kat["fused"]:
        rc, out = lib.decode_vertex_buffer(k["count"], k["size"], bytes.fromhex(k["input"]))
        assert rc == 0
        out = lib.decode_filter(k["filter"], out, k["count"], k["size"])
        assert out.tobytes().hex() == k["expected"], k["name"]


@pytest.mark.parametrize("lib", _libs(), ids=lambda l: l.kind)
def test_version_kat(lib, kat):
    for k in kat["version"]:
        assert lib.decode_vertex_version(bytes.fromhex(k["input"])) == k["rc"]


def test_port_matches_reference_fixtures(port, ref_vectors):
    meta = ref_vectors["codec_meta"]
    for i, (count, vs, version, level) in enumerate(meta):
        enc, dec = ref_vectors[f"codec_{i}_enc"], ref_vectors[f"codec_{i}_dec"]
        rc, out = port.decode_vertex_buffer(int(count), int(vs), enc)
        assert rc == 0 and np.array_equal(out, dec), (i, count, vs, version, level)
    for kind, fname, stride, count in ref_vectors["filter_meta"]:
        stride, count = int(stride), int(count)
        out = port.decode_filter(str(fname), ref_vectors[f"filter_{kind}_in"], count, stride)
        want = ref_vectors[f"filter_{kind}_out"]
        if kind in ("oct8", "color8"):
            # tolerance lanes: the reference uses rsqrtps / rcpps there (vertexfilter.cpp:284,467)
            d = np.abs(out.astype(np.int16) - want.astype(np.int16))
            d = np.minimum(d, 256 - d)
            assert d.max() <= 1, kind
            assert np.array_equal(out[3::4], want[3::4]) or kind == "color8"
        else:
            assert np.array_equal(out, want), kind


def test_memory_safe_truncation(port, kat):
    """every truncated prefix is rejected, the full stream accepted (demo/tests.cpp:521-542),
    one extra byte rejected (:544-557), broken header rejected (:559-572)"""
    for k in kat["codec"][:3]:
        data = np.frombuffer(bytes.fromhex(k["input"]), dtype=np.uint8)
        for cut in range(data.size):
            rc, _ = port.decode_vertex_buffer(k["count"], k["size"], data[:cut])
            assert rc < 0, (k["name"], cut)
        assert port.decode_vertex_buffer(k["count"], k["size"], data)[0] == 0
        assert port.decode_vertex_buffer(k["count"], k["size"], np.concatenate([data, np.zeros(1, np.uint8)]))[0] < 0
        broken = data.copy()
        broken[0] = 0
        assert port.decode_vertex_buffer(k["count"], k["size"], broken)[0] < 0


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_vs_reference_differential(port):
    R = loader.ref()
    rng = np.random.default_rng(7)
    n = 0
    for vs in (4, 8, 12, 16, 24, 32, 48, 64, 256):
        for count in (0, 1, 13, 16, 255, 256, 257, 1031):
            for version, level in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3)):
                kind = n % 4
                if kind == 0:
                    v = rng.integers(0, 256, (count, vs), dtype=np.uint8)
                elif kind == 1:
                    v = np.cumsum(rng.integers(-3, 4, (count, vs)), axis=0).astype(np.uint8)
                elif kind == 2:
                    v = np.cumsum(rng.integers(-300, 300, (count, vs // 4)), axis=0).astype(np.uint32).view(np.uint8).reshape(count, vs)
                else:
                    v = (np.arange(count, dtype=np.uint32)[:, None] << (np.arange(vs // 4, dtype=np.uint32) * 5 % 27)[None, :]).astype(np.uint32).view(np.uint8).reshape(count, vs)
                enc = R.encode_vertex_buffer(v, count, vs, level, version)
                rc_r, out_r = R.decode_vertex_buffer(count, vs, enc)
                rc_p, out_p = port.decode_vertex_buffer(count, vs, enc)
                assert rc_r == 0 and rc_p == 0
                assert np.array_equal(out_r, v.reshape(-1)) and np.array_equal(out_p, out_r), (vs, count, version, level)
                n += 1
    # garbage: identical return codes, identical bytes when accepted (tools/codecfuzz.cpp:143-184 shape)
    for it in range(2000):
        vs = [4, 16, 24, 32][it % 4]
        ln = int(rng.integers(1, 400))
        g = rng.integers(0, 256, ln, dtype=np.uint8)
        g[0] = 0xA0 | (it & 1)
        if it % 3 == 0 and ln > 4:
            g[1 : ln // 2] = rng.integers(0, 4, ln // 2 - 1, dtype=np.uint8)
        a, oa = R.decode_vertex_buffer(66, vs, g)
        b, ob = port.decode_vertex_buffer(66, vs, g)
        assert a == b and (a != 0 or np.array_equal(oa, ob)), it


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_filters_vs_reference():
    from oracle import workloads
    P, R = loader.port(), loader.ref()
    for kind in workloads.C3_KINDS:
        fname, stride, enc = workloads.c3_encoded_elements(kind, 50_003)
        a = R.decode_filter(fname, enc, 50_003, stride)
        b = P.decode_filter(fname, enc, 50_003, stride)
        if kind in ("oct8", "color8"):
            d = np.abs(a.astype(np.int16) - b.astype(np.int16))
            assert np.minimum(d, 256 - d).max() <= 1, kind
        else:
            assert np.array_equal(a, b), kind
