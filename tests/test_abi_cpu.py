"""CPU tests of the boundary: the CUDA library builds for sm_100a, loads, and exports every symbol
that include/meshopt_b200.h declares.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from meshoptimizer_b200 import build
    return build.build()


def test_header_symbols_exported(libpath):
    header = open(os.path.join(ROOT, "include", "meshopt_b200.h")).read()
    declared = set(re.findall(r"MESHOPTIMIZER_API\s+[\w\s\*]+?\b((?:meshopt|mob200)_\w+)\s*\(", header))
    assert {"meshopt_decodeVertexBuffer", "meshopt_decodeVertexVersion", "meshopt_decodeFilterOct", "meshopt_decodeFilterQuat",
            "meshopt_decodeFilterExp", "meshopt_decodeFilterColor", "mob200_plan_run", "mob200_decode_batch_host"} <= declared
    lib = ctypes.CDLL(libpath, mode=os.RTLD_LOCAL)
    for name in declared:
        assert hasattr(lib, name), name
    import meshoptimizer_b200 as mb
    assert set(mb.EXPORTS) == declared


def test_sass_is_sm100a_with_tma(libpath):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lst = subprocess.run([cuobjdump, "-lelf", libpath], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in lst
    sass = subprocess.run([cuobjdump, "-sass", libpath], stdout=subprocess.PIPE, text=True).stdout
    assert "UBLKCP" in sass  # the 1-D TMA bulk copy that stages encoded blocks


def test_version_probe_is_host_logic(libpath, kat):
    """meshopt_decodeVertexVersion is an O(1) header probe; it must work without a device"""
    import meshoptimizer_b200 as mb
    for k in kat["version"]:
        assert mb.decode_vertex_version(bytes.fromhex(k["input"])) == k["rc"]


def test_no_cpu_fallback_without_device(libpath):
    """without a CUDA device the decode entry point must fail loudly, never fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import meshoptimizer_b200 as mb
    data = bytes.fromhex("a0") + bytes(40)
    rc, _ = mb.decode_vertex_buffer_rc(0, 16, data)
    assert rc == mb.ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "meshoptimizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", "").lower() or f == "__init__.py" and "imports\n``oracle/``" in text or \
                    all("import" not in line and "#include" not in line for line in text.splitlines() if "oracle" in line.lower()), f


def test_stream_struct_layout():
    import meshoptimizer_b200 as mb
    assert ctypes.sizeof(mb.Stream) == 48


def test_gltf_scan_rejects_hostile_json():
    """unbounded recursion and unchecked buffer indices / ranges in the glTF scanner"""
    import json

    import meshoptimizer_b200 as mb

    deep = ("[" * 200000).encode()
    with pytest.raises(ValueError):
        mb.gltf_scan(np.frombuffer(b'{"extras":' + deep + b"}", dtype=np.uint8))
    ext = {"buffer": 0, "byteOffset": 0, "byteLength": 40, "byteStride": 4, "count": 10, "mode": "ATTRIBUTES"}
    doc = {
        "buffers": [{"byteLength": 64}, {"byteLength": 40}],
        "bufferViews": [
            {"buffer": 1, "byteOffset": 0, "byteLength": 40, "extensions": {"EXT_meshopt_compression": dict(ext)}},
            {"buffer": 9, "byteOffset": 0, "byteLength": 40, "extensions": {"EXT_meshopt_compression": dict(ext)}},            # no such buffer
            {"buffer": 1, "byteOffset": 8, "byteLength": 40, "extensions": {"EXT_meshopt_compression": dict(ext)}},            # destination beyond buffers[1]
            {"buffer": 1, "byteOffset": 0, "byteLength": 40, "extensions": {"EXT_meshopt_compression": dict(ext, byteOffset=30)}},  # source beyond buffers[0]
            {"buffer": 1, "byteOffset": 0, "byteLength": 2 ** 42, "extensions": {"EXT_meshopt_compression": dict(ext, count=2 ** 40, byteStride=4)}},  # count >= 2^32
            {"buffer": 1, "byteOffset": 0, "byteLength": 40, "extensions": {"EXT_meshopt_compression": dict(ext, buffer=5)}},   # no such source buffer
        ],
    }
    views, sizes, info = mb.gltf_scan(np.frombuffer(json.dumps(doc).encode(), dtype=np.uint8))
    assert info.view_count == 6 and sizes == [64, 40]
    assert [views[i].status for i in range(6)] == [0, mb.ERR_ARGUMENT, mb.ERR_ARGUMENT, mb.ERR_ARGUMENT, mb.ERR_ARGUMENT, mb.ERR_ARGUMENT]
    assert info.invalid_views == 5
