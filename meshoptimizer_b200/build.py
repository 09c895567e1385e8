"""In-tree build of the CUDA library (sm_100a only): nvcc -> meshoptimizer_b200/lib/libmeshopt_b200.so.

The .so is git-ignored but travels to the GPU box with the repository snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmeshopt_b200.so")
SOURCES = ["mob200_kernels.cu", "mob200_api.cu", "mob200_index.cu", "mob200_meshlet.cu", "mob200_gltf.cpp", "mob200_encode.cpp"]
HEADERS = ["mob200_common.h", "mob200_host.h", "mob200_kernels.h", "mob200_filters.cuh", "mob200_device.cuh", "mob200_walker.cuh", "mob200_walker_wide.cuh", "mob200_decoder.cuh", os.path.join("..", "..", "include", "meshopt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # decode filters: every FP op separately rounded
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines) -> str:
    """Diagnostics: the same sources with extra -D switches -> lib/variants/<name>.so (tools/ab.sh)."""
    out = os.path.join(LIB_DIR, "variants", name + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    debug = ["-DMOB200_DEBUG_COUNTERS"] if os.environ.get("MOB200_DEBUG_COUNTERS") == "1" else []
    cmd = [_nvcc()] + NVCC_FLAGS + debug + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
