"""Produce the glTF fixtures of the bufferView front-end tests (tests/golden/gltf_*.glb + gltf_expected.npz).

Run in the build container (needs /root/reference):  python tests/golden/make_gltf.py
  1. `make -C oracle gltfpack ref` builds the UNMODIFIED reference gltfpack and library (oracle/_ref/);
  2. a synthetic scene (two wavy grid meshes with normals and texture coordinates, one point cloud) is written as
     Wavefront OBJ and packed with the reference gltfpack in several compression settings
     (-c, -cc, -cc with float positions/normals so that the exponential / octahedral filters appear, -ce khr);
  3. for every packed .glb the expected decompressed buffer is produced by the REFERENCE decoders
     (oracle/_ref/libmeshopt_ref.so), view by view, exactly as gltf/parsegltf.cpp:561-627 does.
The .glb files and the expected buffers are committed; the GPU tests never need /root/reference.
"""
import json, os, struct, subprocess, sys, tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

GLTFPACK = os.path.join(ROOT, "oracle", "_ref", "gltfpack")
VARIANTS = {
    "c": ["-c"],
    "cc": ["-cc"],
    "cc_float": ["-cc", "-vpf", "-vnf", "-vtf"],
    "khr": ["-cc", "-ce", "khr"],
}
FILTERS = {"NONE": None, "OCTAHEDRAL": "oct", "QUATERNION": "quat", "EXPONENTIAL": "exp", "COLOR": "color"}


def write_obj(path, side=28):
    with open(path, "w") as f:
        for m in range(2):
            f.write(f"o grid{m}\n")
            for y in range(side + 1):
                for x in range(side + 1):
                    u, v = x / side, y / side
                    z = 0.15 * np.sin(6.0 * u + m) * np.cos(5.0 * v)
                    f.write(f"v {u * 2 - 1 + 3 * m:.6f} {v * 2 - 1:.6f} {z:.6f}\n")
            for y in range(side + 1):
                for x in range(side + 1):
                    u, v = x / side, y / side
                    dx = -0.9 * np.cos(6.0 * u + m) * np.cos(5.0 * v)
                    dy = 0.75 * np.sin(6.0 * u + m) * np.sin(5.0 * v)
                    n = np.array([dx, dy, 1.0])
                    n /= np.linalg.norm(n)
                    f.write(f"vn {n[0]:.6f} {n[1]:.6f} {n[2]:.6f}\n")
                    f.write(f"vt {u:.6f} {v:.6f}\n")
            base = m * (side + 1) * (side + 1)
            for y in range(side):
                for x in range(side):
                    a = base + y * (side + 1) + x + 1
                    b, c, d = a + 1, a + side + 1, a + side + 2
                    f.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
                    f.write(f"f {b}/{b}/{b} {d}/{d}/{d} {c}/{c}/{c}\n")


def split_glb(blob):
    assert blob[:4] == b"glTF"
    at, chunks = 12, {}
    while at + 8 <= len(blob):
        n, t = struct.unpack_from("<II", blob, at)
        chunks[t] = blob[at + 8 : at + 8 + n]
        at += 8 + ((n + 3) & ~3)
    return json.loads(chunks[0x4E4F534A]), chunks.get(0x004E4942, b"")


def expected_buffers(blob):
    """decompress every compressed view with the reference decoders (the loop of parsegltf.cpp:561-627)"""
    R = loader.ref()
    doc, bin_chunk = split_glb(blob)
    out = {}
    summary = []
    for i, view in enumerate(doc.get("bufferViews", [])):
        ext = view.get("extensions", {})
        mc = ext.get("EXT_meshopt_compression") or ext.get("KHR_meshopt_compression")
        if not mc:
            continue
        assert mc["buffer"] == 0
        src = np.frombuffer(bin_chunk, np.uint8)[mc.get("byteOffset", 0) : mc.get("byteOffset", 0) + mc["byteLength"]]
        count, stride, mode = mc["count"], mc["byteStride"], mc["mode"]
        if mode == "ATTRIBUTES":
            rc, dec = R.decode_vertex_buffer(count, stride, src)
        elif mode == "TRIANGLES":
            rc, dec = R.decode_index("triangles", count, stride, src)
        else:
            rc, dec = R.decode_index("sequence", count, stride, src)
        assert rc == 0, (i, mode, rc)
        dec = np.ascontiguousarray(dec).view(np.uint8).reshape(-1)
        filt = FILTERS[mc.get("filter", "NONE")]
        if filt:
            dec = R.decode_filter(filt, dec, count, stride)
        dst = out.setdefault(view["buffer"], np.zeros(doc["buffers"][view["buffer"]]["byteLength"], np.uint8))
        dst[view.get("byteOffset", 0) : view.get("byteOffset", 0) + view["byteLength"]] = dec[: view["byteLength"]]
        summary.append((mode, mc.get("filter", "NONE"), count, stride))
    return out, summary


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "gltfpack", "ref"])
    expected = {}
    with tempfile.TemporaryDirectory() as tmp:
        obj = os.path.join(tmp, "scene.obj")
        write_obj(obj)
        for name, flags in VARIANTS.items():
            glb = os.path.join(HERE, f"gltf_{name}.glb")
            subprocess.check_call([GLTFPACK] + flags + ["-i", obj, "-o", glb], stdout=subprocess.DEVNULL)
            blob = open(glb, "rb").read()
            bufs, summary = expected_buffers(blob)
            for b, data in bufs.items():
                expected[f"{name}_buffer{b}"] = data
            print(name, len(blob), "bytes;", len(summary), "compressed views:", sorted(set(summary)))
    np.savez_compressed(os.path.join(HERE, "gltf_expected.npz"), **expected)


if __name__ == "__main__":
    main()
