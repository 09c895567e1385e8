"""CPU tests of the meshlet checker (oracle/meshletcodec_oracle.c) against the unmodified reference decoder
(oracle/_ref) on encoder-produced, truncated and corrupted meshlets, in all four output formats."""
import numpy as np
import pytest

from tests.meshlet_cases import corruptions, meshlets


def test_port_matches_reference(port):
    from oracle import loader
    if not loader.have_ref():
        pytest.skip("oracle/_ref not built")
    R = loader.ref()
    cases = 0
    for ci, (name, verts, tris) in enumerate(meshlets()):
        enc = R.encode_meshlet(verts, tris)
        for vs in (2, 4):
            for ts in (3, 4):
                a = R.decode_meshlet(verts.size, vs, tris.shape[0], ts, enc)
                b = port.decode_meshlet(verts.size, vs, tris.shape[0], ts, enc)
                assert a[0] == b[0] == 0, (name, vs, ts, a[0], b[0])
                assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), (name, vs, ts)
                assert np.array_equal(a[1].astype(np.uint32), verts if vs == 4 else verts & 0xFFFF)
                if ts == 3:  # triangles survive up to rotation (the encoder may rotate them)
                    t, w = a[2].astype(np.int32), tris.astype(np.int32)
                    assert ((t == w).all(1) | (t == np.roll(w, 1, 1)).all(1) | (t == np.roll(w, 2, 1)).all(1)).all(), name
                for e in corruptions(enc, seed=ci):
                    a = R.decode_meshlet(verts.size, vs, tris.shape[0], ts, e)
                    b = port.decode_meshlet(verts.size, vs, tris.shape[0], ts, e)
                    assert a[0] == b[0], (name, vs, ts, e.size, a[0], b[0])
                    if a[0] == 0:
                        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
                    cases += 1
    assert cases > 500
