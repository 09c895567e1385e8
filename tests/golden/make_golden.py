"""Generate tests/golden/ref_vectors.npz by running the UNMODIFIED reference (oracle/_ref) in the build
container:  python tests/golden/make_golden.py

Contents (all produced by reference code: meshopt_encodeVertexBufferLevel, meshopt_decodeVertexBuffer,
meshopt_encodeFilter*, meshopt_decodeFilter*):
  codec_{i}_enc / codec_{i}_dec + codec_meta[i] = (vertex_count, vertex_size, version, level)
  filter_{kind}_in / filter_{kind}_out + filter_meta (kind, filter name, stride, count)
The fixtures pin the C restatement (CPU tests) and the CUDA path (GPU tests) to reference OUTPUTS even
where /root/reference and oracle/_ref are absent.
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import loader, workloads  # noqa: E402


def vertex_data(kind, count, vs, rng):
    if kind == 0:  # smooth bytes
        return np.cumsum(rng.integers(-3, 4, (count, vs)), axis=0).astype(np.uint8)
    if kind == 1:  # 16-bit lanes crossing byte boundaries (exercises channel mode 1 at level >= 2)
        w = (0xf0 + np.cumsum(rng.integers(-5, 9, (count, vs // 2)), axis=0)).astype(np.uint16)
        return w.view(np.uint8).reshape(count, vs)
    if kind == 2:  # 32-bit lanes with shifted bit fields (exercises xor/rotate at level 3)
        i = np.arange(count, dtype=np.uint64)[:, None]
        sh = (np.arange(vs // 4, dtype=np.uint64) * 7 % 29)[None, :]
        w = ((i * 3 + rng.integers(0, 3, (count, vs // 4)).astype(np.uint64)) << sh).astype(np.uint32)
        return w.view(np.uint8).reshape(count, vs)
    return rng.integers(0, 256, (count, vs), dtype=np.uint8)  # noise: literal / 8-bit groups, many sentinels


def main():
    R = loader.ref()
    rng = np.random.default_rng(20261017)
    out, meta = {}, []
    i = 0
    for vs in (4, 12, 16, 32, 64, 256):
        for count in (1, 13, 257):
            for version, level in ((0, 0), (1, 2), (1, 3)):
                v = vertex_data(i % 4, count, vs, rng)
                enc = R.encode_vertex_buffer(v, count, vs, level, version)
                rc, dec = R.decode_vertex_buffer(count, vs, enc)
                assert rc == 0 and np.array_equal(dec, v.reshape(-1))
                out[f"codec_{i}_enc"], out[f"codec_{i}_dec"] = enc, dec
                meta.append((count, vs, version, level))
                i += 1
    out["codec_meta"] = np.array(meta, dtype=np.int64)

    fmeta = []
    for kind in workloads.C3_KINDS:
        count = 1024
        fname, stride, enc = workloads.c3_encoded_elements(kind, count)
        dec = R.decode_filter(fname, enc, count, stride)
        out[f"filter_{kind}_in"], out[f"filter_{kind}_out"] = enc, dec
        fmeta.append((kind, fname, stride, count))
    out["filter_meta"] = np.array(fmeta, dtype="U16")
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", i, "codec streams;", len(fmeta), "filter sets")


if __name__ == "__main__":
    main()
