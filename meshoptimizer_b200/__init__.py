"""meshoptimizer_b200 -- B200-native (sm_100a) vertex-buffer decode for meshoptimizer streams.

Python host mirror of the reference interface for this path, on top of the C ABI declared in
``include/meshopt_b200.h`` (``lib/libmeshopt_b200.so``).  Names and argument meaning follow the
reference (``src/meshoptimizer.h:396-424``; JS wrapper ``js/meshopt_decoder.mjs:164-193``):

* ``decode_vertex_buffer(count, size, source, filter=None)``  -> host bytes in, numpy bytes out
* ``decode_vertex_version(source)``
* ``decode_filter_oct / quat / exp / color(buffer, count, stride)``  (in place, host numpy)
* ``Context`` / ``Plan``: the batched device-pointer variant (torch tensors only carry device memory)

There is NO CPU fallback: every entry point goes through the CUDA library and raises
``RuntimeError`` when the library is missing or no CUDA device is usable.  Nothing here imports
``oracle/``.
"""
from __future__ import annotations

import ctypes
import weakref
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p
from typing import Iterable, Optional, Sequence

import numpy as np

from . import build as _build

FILTER_NONE, FILTER_OCTAHEDRAL, FILTER_QUATERNION, FILTER_EXPONENTIAL, FILTER_COLOR = 0, 1, 2, 3, 4
_FILTER_BY_NAME = {
    None: 0, "NONE": 0, "none": 0,
    "OCTAHEDRAL": 1, "oct": 1,
    "QUATERNION": 2, "quat": 2,
    "EXPONENTIAL": 3, "exp": 3,
    "COLOR": 4, "color": 4,
}
ERR_CUDA, ERR_ARGUMENT, ERR_SIDECAR = -100, -101, -102
RUN_BLOCK_PARALLEL = 1


INDEX_TRIANGLES, INDEX_SEQUENCE = 0, 1
GLTF_ATTRIBUTES, GLTF_TRIANGLES, GLTF_INDICES = 0, 1, 2


class IndexStream(ctypes.Structure):
    """mirror of ``mob200_IndexStream``"""
    _fields_ = [
        ("src", c_void_p),
        ("src_size", c_size_t),
        ("dst", c_void_p),
        ("index_count", c_size_t),
        ("index_size", c_size_t),
        ("kind", c_int),
        ("status", c_int),
    ]


class Meshlet(ctypes.Structure):
    """mirror of ``mob200_Meshlet``"""
    _fields_ = [
        ("src", c_void_p), ("src_size", c_size_t),
        ("vertices", c_void_p), ("vertex_count", c_size_t), ("vertex_size", c_size_t),
        ("triangles", c_void_p), ("triangle_count", c_size_t), ("triangle_size", c_size_t),
        ("status", c_int),
    ]


class GltfView(ctypes.Structure):
    """mirror of ``mob200_GltfView``"""
    _fields_ = [
        ("view", c_size_t),
        ("mode", c_int),
        ("filter", c_int),
        ("src_buffer", c_size_t), ("src_offset", c_size_t), ("src_size", c_size_t),
        ("count", c_size_t), ("stride", c_size_t),
        ("dst_buffer", c_size_t), ("dst_offset", c_size_t), ("dst_size", c_size_t),
        ("status", c_int),
    ]


class GltfInfo(ctypes.Structure):
    """mirror of ``mob200_GltfInfo``"""
    _fields_ = [
        ("json_offset", c_size_t), ("json_size", c_size_t),
        ("bin_offset", c_size_t), ("bin_size", c_size_t),
        ("buffer_count", c_size_t), ("view_count", c_size_t), ("invalid_views", c_size_t),
    ]


class Segment(ctypes.Structure):
    """mirror of ``mob200_Segment``"""
    _fields_ = [
        ("first_vertex", c_size_t), ("vertex_count", c_size_t),
        ("offset", c_size_t), ("size", c_size_t),
        ("sidecar_offset", c_size_t), ("sidecar_entries", c_size_t),
    ]


class Stream(ctypes.Structure):
    """mirror of ``mob200_Stream``"""
    _fields_ = [
        ("src", c_void_p),
        ("src_size", c_size_t),
        ("dst", c_void_p),
        ("vertex_count", c_size_t),
        ("vertex_size", c_size_t),
        ("filter", c_int),
        ("status", c_int),
    ]


_LIB = None

EXPORTS = [
    "meshopt_decodeVertexBuffer", "meshopt_decodeVertexVersion",
    "meshopt_decodeFilterOct", "meshopt_decodeFilterQuat", "meshopt_decodeFilterExp", "meshopt_decodeFilterColor",
    "mob200_context_create", "mob200_context_destroy", "mob200_plan_create", "mob200_plan_destroy",
    "mob200_plan_run", "mob200_plan_status", "mob200_plan_launches", "mob200_decode_batch_device",
    "mob200_decode_batch_host", "mob200_decode_batch_host_sidecar", "mob200_filter_device", "mob200_context_sm_count", "mob200_version",
    "mob200_plan_last_timing", "mob200_plan_timing_history", "mob200_plan_debug_counters", "mob200_plan_create_ms",
    "mob200_encode_vertex_bound", "mob200_encode_vertex_buffer", "mob200_segment_count", "mob200_encode_segments_bound", "mob200_segments_sidecar_entries", "mob200_encode_segments",
    "mob200_shard_streams", "mob200_decode_batch_multi_host",
    "mob200_sidecar_entries", "mob200_plan_create_sidecar", "mob200_plan_run_ex", "mob200_plan_has_offsets", "mob200_plan_export_sidecar",
    "meshopt_decodeIndexBuffer", "meshopt_decodeIndexVersion", "meshopt_decodeIndexSequence",
    "mob200_decode_index_batch_device", "mob200_decode_index_batch_host",
    "mob200_gltf_scan", "mob200_gltf_decode_host", "mob200_gltf_decode_device",
    "mob200_debug_last_kernel_ms", "meshopt_decodeMeshlet", "meshopt_decodeMeshletRaw", "mob200_decode_meshlet_batch_device", "mob200_decode_meshlet_batch_host",
]


def library_path() -> str:
    # MOB200_LIB: a variant build of the same sources (tools/ab.sh, diagnostics only)
    return os.environ.get("MOB200_LIB") or _build.LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the CUDA library (never builds implicitly on a GPU box: the .so ships in-tree)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m meshoptimizer_b200.build` (there is no CPU fallback)")
    L = ctypes.CDLL(path, mode=os.RTLD_LOCAL)
    L.meshopt_decodeVertexBuffer.restype = c_int
    L.meshopt_decodeVertexBuffer.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
    L.meshopt_decodeVertexVersion.restype = c_int
    L.meshopt_decodeVertexVersion.argtypes = [c_void_p, c_size_t]
    for name in ("Oct", "Quat", "Exp", "Color"):
        f = getattr(L, "meshopt_decodeFilter" + name)
        f.restype = None
        f.argtypes = [c_void_p, c_size_t, c_size_t]
    L.mob200_context_create.restype = c_int
    L.mob200_context_create.argtypes = [POINTER(c_void_p), c_int]
    L.mob200_context_destroy.restype = None
    L.mob200_context_destroy.argtypes = [c_void_p]
    L.mob200_plan_create.restype = c_int
    L.mob200_plan_create.argtypes = [c_void_p, POINTER(Stream), c_size_t, POINTER(c_void_p)]
    L.mob200_plan_destroy.restype = None
    L.mob200_plan_destroy.argtypes = [c_void_p]
    L.mob200_plan_run.restype = c_int
    L.mob200_plan_run.argtypes = [c_void_p, c_void_p]
    L.mob200_plan_status.restype = c_int
    L.mob200_plan_status.argtypes = [c_void_p, POINTER(c_int), c_void_p]
    L.mob200_plan_launches.restype = c_int
    L.mob200_plan_launches.argtypes = [c_void_p]
    L.mob200_plan_last_timing.restype = c_int
    L.mob200_plan_last_timing.argtypes = [c_void_p, POINTER(c_float)]
    L.mob200_plan_create_ms.restype = c_float
    L.mob200_plan_create_ms.argtypes = [c_void_p]
    L.mob200_encode_vertex_bound.restype = c_size_t
    L.mob200_encode_vertex_bound.argtypes = [c_size_t, c_size_t]
    L.mob200_encode_vertex_buffer.restype = c_size_t
    L.mob200_encode_vertex_buffer.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_int, c_void_p]
    L.mob200_segment_count.restype = c_size_t
    L.mob200_segment_count.argtypes = [c_size_t, c_size_t]
    L.mob200_encode_segments_bound.restype = c_size_t
    L.mob200_encode_segments_bound.argtypes = [c_size_t, c_size_t, c_size_t]
    L.mob200_segments_sidecar_entries.restype = c_size_t
    L.mob200_segments_sidecar_entries.argtypes = [c_size_t, c_size_t, c_size_t]
    L.mob200_encode_segments.restype = c_int
    L.mob200_encode_segments.argtypes = [c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, POINTER(c_size_t), POINTER(Segment), c_size_t, c_void_p, c_size_t]
    L.mob200_shard_streams.restype = c_int
    L.mob200_shard_streams.argtypes = [POINTER(c_size_t), c_size_t, c_int, POINTER(c_int)]
    L.mob200_decode_batch_multi_host.restype = c_int
    L.mob200_decode_batch_multi_host.argtypes = [POINTER(c_int), c_int, POINTER(Stream), c_size_t, POINTER(c_void_p), POINTER(c_float)]
    L.mob200_sidecar_entries.restype = c_size_t
    L.mob200_sidecar_entries.argtypes = [c_size_t, c_size_t]
    L.mob200_plan_create_sidecar.restype = c_int
    L.mob200_plan_create_sidecar.argtypes = [c_void_p, POINTER(Stream), c_size_t, POINTER(c_void_p), POINTER(c_void_p)]
    L.mob200_plan_run_ex.restype = c_int
    L.mob200_plan_run_ex.argtypes = [c_void_p, c_void_p, c_int]
    L.mob200_plan_has_offsets.restype = c_int
    L.mob200_plan_has_offsets.argtypes = [c_void_p]
    L.mob200_plan_export_sidecar.restype = c_int
    L.mob200_plan_export_sidecar.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]
    L.mob200_plan_debug_counters.restype = c_int
    L.mob200_plan_debug_counters.argtypes = [c_void_p, POINTER(ctypes.c_ulonglong), c_int, c_int]
    L.mob200_plan_timing_history.restype = c_int
    L.mob200_plan_timing_history.argtypes = [c_void_p, c_int, POINTER(c_float)]
    L.mob200_decode_batch_device.restype = c_int
    L.mob200_decode_batch_device.argtypes = [c_void_p, POINTER(Stream), c_size_t, c_void_p]
    L.mob200_decode_batch_host.restype = c_int
    L.mob200_decode_batch_host.argtypes = [c_void_p, POINTER(Stream), c_size_t]
    L.mob200_filter_device.restype = c_int
    L.mob200_filter_device.argtypes = [c_int, c_void_p, c_size_t, c_size_t, c_void_p]
    L.mob200_context_sm_count.restype = c_int
    L.mob200_context_sm_count.argtypes = [c_void_p]
    L.mob200_version.restype = c_char_p
    for name in ("meshopt_decodeIndexBuffer", "meshopt_decodeIndexSequence"):
        f = getattr(L, name)
        f.restype = c_int
        f.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
    L.meshopt_decodeIndexVersion.restype = c_int
    L.meshopt_decodeIndexVersion.argtypes = [c_void_p, c_size_t]
    L.mob200_decode_index_batch_device.restype = c_int
    L.mob200_decode_index_batch_device.argtypes = [c_void_p, POINTER(IndexStream), c_size_t, c_void_p]
    L.mob200_decode_index_batch_host.restype = c_int
    L.mob200_decode_index_batch_host.argtypes = [c_void_p, POINTER(IndexStream), c_size_t]
    L.meshopt_decodeMeshlet.restype = c_int
    L.meshopt_decodeMeshlet.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
    L.meshopt_decodeMeshletRaw.restype = c_int
    L.meshopt_decodeMeshletRaw.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t]
    L.mob200_debug_last_kernel_ms.restype = c_float
    L.mob200_decode_meshlet_batch_device.restype = c_int
    L.mob200_decode_meshlet_batch_device.argtypes = [c_void_p, POINTER(Meshlet), c_size_t, c_void_p]
    L.mob200_decode_meshlet_batch_host.restype = c_int
    L.mob200_decode_meshlet_batch_host.argtypes = [c_void_p, POINTER(Meshlet), c_size_t]
    L.mob200_gltf_scan.restype = c_int
    L.mob200_gltf_scan.argtypes = [c_void_p, c_size_t, POINTER(GltfView), c_size_t, POINTER(c_size_t), c_size_t, POINTER(GltfInfo)]
    L.mob200_gltf_decode_host.restype = c_int
    L.mob200_gltf_decode_host.argtypes = [c_void_p, POINTER(GltfView), c_size_t, c_size_t, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p), POINTER(c_size_t)]
    L.mob200_gltf_decode_device.restype = c_int
    L.mob200_gltf_decode_device.argtypes = [c_void_p, POINTER(GltfView), c_size_t, c_size_t, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p), POINTER(c_size_t), c_void_p]
    L.mob200_decode_batch_host_sidecar.restype = c_int
    L.mob200_decode_batch_host_sidecar.argtypes = [c_void_p, POINTER(Stream), c_size_t, POINTER(c_void_p)]
    _LIB = L
    return L


def _as_u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        return np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def _filter_id(f) -> int:
    if isinstance(f, int):
        return f
    return _FILTER_BY_NAME[f]


def _check_vertex_size(vertex_size: int) -> None:
    # the reference asserts these (src/vertexcodec.cpp:1803-1804)
    if not (0 < vertex_size <= 256 and vertex_size % 4 == 0):
        raise ValueError("vertex_size must be a multiple of 4 in (0, 256]")


# ---------------------------------------------------------------------------------------------
# drop-in, host memory (synchronous)
# ---------------------------------------------------------------------------------------------

def decode_vertex_version(source) -> int:
    """``meshopt_decodeVertexVersion``: 0 or 1, -1 for an invalid header."""
    src = _as_u8(source)
    return int(lib().meshopt_decodeVertexVersion(src.ctypes.data if src.size else None, src.size))


def decode_vertex_buffer_rc(count: int, size: int, source, target: Optional[np.ndarray] = None):
    """``meshopt_decodeVertexBuffer`` on host memory; returns (return code, decoded uint8 array)."""
    _check_vertex_size(size)
    src = _as_u8(source)
    out = target if target is not None else np.zeros(count * size, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= count * size and out.flags.c_contiguous
    rc = lib().meshopt_decodeVertexBuffer(out.ctypes.data if count else None, count, size, src.ctypes.data if src.size else None, src.size)
    return int(rc), out[: count * size]


def decode_vertex_buffer(count: int, size: int, source, filter=None) -> np.ndarray:
    """Decode one stream and optionally apply a decode filter (as ``MeshoptDecoder.decodeVertexBuffer``
    does, js/meshopt_decoder.mjs:48-66).  Raises on a non-zero return code."""
    fid = _filter_id(filter)
    if fid == FILTER_NONE:
        rc, out = decode_vertex_buffer_rc(count, size, source)
    else:
        outs, rcs = decode_batch_host([(source, count, size, fid)])
        rc, out = rcs[0], outs[0]
    if rc != 0:
        raise RuntimeError(f"Malformed buffer data: {rc}")
    return out


def _decode_filter(name: str, buffer: np.ndarray, count: int, stride: int) -> np.ndarray:
    buf = buffer.view(np.uint8).reshape(-1)
    assert buf.flags.c_contiguous and buf.flags.writeable and buf.size >= count * stride
    getattr(lib(), "meshopt_decodeFilter" + name)(buf.ctypes.data, count, stride)
    return buffer


def decode_filter_oct(buffer: np.ndarray, count: int, stride: int) -> np.ndarray:
    if stride not in (4, 8):
        raise ValueError("stride must be 4 or 8")
    return _decode_filter("Oct", buffer, count, stride)


def decode_filter_quat(buffer: np.ndarray, count: int, stride: int) -> np.ndarray:
    if stride != 8:
        raise ValueError("stride must be 8")
    return _decode_filter("Quat", buffer, count, stride)


def decode_filter_exp(buffer: np.ndarray, count: int, stride: int) -> np.ndarray:
    if stride <= 0 or stride % 4:
        raise ValueError("stride must be a positive multiple of 4")
    return _decode_filter("Exp", buffer, count, stride)


def decode_filter_color(buffer: np.ndarray, count: int, stride: int) -> np.ndarray:
    if stride not in (4, 8):
        raise ValueError("stride must be 4 or 8")
    return _decode_filter("Color", buffer, count, stride)


# ---------------------------------------------------------------------------------------------
# batched variants
# ---------------------------------------------------------------------------------------------

class Context:
    """``mob200_Context``: per-device scratch, staging buffers and a private CUDA stream."""

    def __init__(self, device: int = -1):
        h = c_void_p()
        rc = lib().mob200_context_create(ctypes.byref(h), device)
        if rc != 0:
            raise RuntimeError(f"mob200_context_create failed ({rc}): no usable CUDA device? (there is no CPU fallback)")
        self.handle = h

    @property
    def sm_count(self) -> int:
        return int(lib().mob200_context_sm_count(self.handle))

    def close(self) -> None:
        if self.handle:
            # plans point at their context (include/meshopt_b200.h: destroy a context's plans first)
            for p in list(getattr(self, "_plans", ())):
                p.close()
            lib().mob200_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_CTX: Optional[Context] = None


def default_context() -> Context:
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context(-1)
    return _DEFAULT_CTX


def make_streams(items: Sequence[tuple]) -> ctypes.Array:
    """items: (src_ptr, src_size, dst_ptr, vertex_count, vertex_size, filter) with integer addresses."""
    arr = (Stream * len(items))()
    for i, (src, src_size, dst, count, vs, filt) in enumerate(items):
        arr[i].src = src
        arr[i].src_size = src_size
        arr[i].dst = dst
        arr[i].vertex_count = count
        arr[i].vertex_size = vs
        arr[i].filter = _filter_id(filt)
        arr[i].status = 0
    return arr


class Plan:
    """``mob200_Plan``: a prepared batch of device-resident streams that can be run repeatedly.

    ``sidecars``: optional list with one uint32 array (``sidecar_entries(count, size)`` block offsets) or ``None``
    per stream; with them ``run(block_parallel=True)`` walks every block on its own GPU lane."""

    def __init__(self, ctx: Context, streams: ctypes.Array, sidecars: Optional[Sequence] = None):
        self.ctx = ctx
        self.n = len(streams)
        h = c_void_p()
        if sidecars is None:
            rc = lib().mob200_plan_create(ctx.handle, streams, self.n, ctypes.byref(h))
        else:
            assert len(sidecars) == self.n
            keep = [None if sc is None else np.ascontiguousarray(sc, dtype=np.uint32) for sc in sidecars]
            ptrs = (c_void_p * max(self.n, 1))()
            for i, sc in enumerate(keep):
                if sc is not None:
                    assert sc.size == sidecar_entries(streams[i].vertex_count, streams[i].vertex_size), "sidecar length"
                    ptrs[i] = sc.ctypes.data
            rc = lib().mob200_plan_create_sidecar(ctx.handle, streams, self.n, ptrs, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"mob200_plan_create failed ({rc})")
        self.handle = h
        if not hasattr(ctx, "_plans"):
            ctx._plans = weakref.WeakSet()
        ctx._plans.add(self)

    def run(self, cuda_stream: int = 0, block_parallel: bool = False) -> None:
        rc = lib().mob200_plan_run_ex(self.handle, c_void_p(cuda_stream), RUN_BLOCK_PARALLEL if block_parallel else 0)
        if rc != 0:
            raise RuntimeError(f"mob200_plan_run failed ({rc})")

    @property
    def has_offsets(self) -> bool:
        return bool(lib().mob200_plan_has_offsets(self.handle))

    def export_sidecar(self, stream_index: int, vertex_count: int, vertex_size: int, cuda_stream: int = 0) -> np.ndarray:
        """block offsets of one stream after a run of the serial walk (raises if that stream's walk failed)"""
        out = np.zeros(max(1, sidecar_entries(vertex_count, vertex_size)), dtype=np.uint32)
        rc = lib().mob200_plan_export_sidecar(self.handle, stream_index, out.ctypes.data, out.size, c_void_p(cuda_stream))
        if rc < 0:
            raise RuntimeError(f"mob200_plan_export_sidecar failed ({rc})")
        return out[:rc]

    def status(self, cuda_stream: int = 0) -> np.ndarray:
        st = np.zeros(max(self.n, 1), dtype=np.int32)
        rc = lib().mob200_plan_status(self.handle, st.ctypes.data_as(POINTER(c_int)), c_void_p(cuda_stream))
        if rc < 0:
            raise RuntimeError(f"mob200_plan_status failed ({rc})")
        return st[: self.n]

    @property
    def launches(self) -> int:
        return int(lib().mob200_plan_launches(self.handle))

    @property
    def create_ms(self) -> float:
        return float(lib().mob200_plan_create_ms(self.handle))

    def last_timing(self) -> float:
        """milliseconds of the most recent run's kernel (CUDA events on the launching stream)"""
        a = c_float()
        rc = lib().mob200_plan_last_timing(self.handle, ctypes.byref(a))
        if rc != 0:
            raise RuntimeError(f"mob200_plan_last_timing failed ({rc})")
        return a.value

    def debug_counters(self, reset: bool = True):
        """cycle counters accumulated by the kernel (see mob200_plan_debug_counters)"""
        out = (ctypes.c_ulonglong * 16)()
        rc = lib().mob200_plan_debug_counters(self.handle, out, 16, int(reset))
        if rc != 0:
            raise RuntimeError(f"mob200_plan_debug_counters failed ({rc})")
        names = ["decoder_total", "decoder_wait_full", "decoder_wait_carry", "decoder_wait_tile", "producer_total", "producer_meta", "producer_wait_slot", "producer_lookback",
                 "walker_total", "walker_wait_prev", "walker_wait_hard", "walker_refills", "walker_hard_waits", "r13", "r14", "r15"]
        return dict(zip(names, [int(v) for v in out]))

    def timing_history(self, max_runs: int = 64):
        """kernel durations (ms) of the most recent runs, oldest first"""
        a = (c_float * max_runs)()
        n = lib().mob200_plan_timing_history(self.handle, max_runs, a)
        if n < 0:
            raise RuntimeError(f"mob200_plan_timing_history failed ({n})")
        return [float(a[i]) for i in range(n)]

    def close(self) -> None:
        if self.handle:
            lib().mob200_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sidecar_entries(vertex_count: int, vertex_size: int) -> int:
    return int(lib().mob200_sidecar_entries(vertex_count, vertex_size))


# ---------------------------------------------------------------------------------------------
# encoder-side helper: the segmenter (host code, no CUDA; reference src/vertexcodec.cpp:1615-1693)
# ---------------------------------------------------------------------------------------------

def encode_vertex_buffer(vertices, vertex_count: int, vertex_size: int, level: int = 2, version: int = 1, with_sidecar: bool = False):
    """``mob200_encode_vertex_buffer``: one reference-format stream -> uint8 array (and its block offsets)"""
    _check_vertex_size(vertex_size)
    v = _as_u8(vertices)
    assert v.size >= vertex_count * vertex_size
    buf = np.empty(int(lib().mob200_encode_vertex_bound(vertex_count, vertex_size)), dtype=np.uint8)
    side = np.zeros(max(1, sidecar_entries(vertex_count, vertex_size)), dtype=np.uint32) if with_sidecar else None
    n = lib().mob200_encode_vertex_buffer(buf.ctypes.data, buf.size, v.ctypes.data if v.size else None, vertex_count, vertex_size, level, version,
                                          side.ctypes.data if with_sidecar else None)
    if n == 0:
        raise RuntimeError("mob200_encode_vertex_buffer failed")
    out = buf[:n].copy()
    return (out, side[: sidecar_entries(vertex_count, vertex_size)]) if with_sidecar else out


def encode_segments(vertices, vertex_count: int, vertex_size: int, segment_vertices: int, level: int = 2, version: int = 1, threads: int = 0, with_sidecars: bool = True):
    """``mob200_encode_segments``: (blob uint8, list of Segment, sidecar uint32 array or None)"""
    _check_vertex_size(vertex_size)
    v = _as_u8(vertices)
    assert v.size >= vertex_count * vertex_size
    L = lib()
    n = int(L.mob200_segment_count(vertex_count, segment_vertices))
    blob = np.empty(int(L.mob200_encode_segments_bound(vertex_count, vertex_size, segment_vertices)) + 16, dtype=np.uint8)
    segs = (Segment * max(n, 1))()
    side = np.zeros(max(1, int(L.mob200_segments_sidecar_entries(vertex_count, vertex_size, segment_vertices))), dtype=np.uint32) if with_sidecars else None
    used = c_size_t(0)
    rc = L.mob200_encode_segments(v.ctypes.data if v.size else None, vertex_count, vertex_size, segment_vertices, level, version, threads,
                                  blob.ctypes.data, blob.size, ctypes.byref(used), segs, n, side.ctypes.data if with_sidecars else None, side.size if with_sidecars else 0)
    if rc < 0:
        raise RuntimeError(f"mob200_encode_segments failed ({rc})")
    return blob[: used.value], [segs[i] for i in range(n)], side


def decode_batch_host(items: Iterable[tuple], ctx: Optional[Context] = None):
    """items: (source bytes/array, vertex_count, vertex_size, filter).  Host memory in, host memory
    out through ``mob200_decode_batch_host``.  Returns (list of uint8 arrays, list of return codes)."""
    items = list(items)
    ctx = ctx or default_context()
    srcs = [_as_u8(it[0]) for it in items]
    outs = [np.zeros(max(it[1] * it[2], 1), dtype=np.uint8) for it in items]
    for it in items:
        _check_vertex_size(it[2])
    arr = make_streams([
        (s.ctypes.data if s.size else None, s.size, o.ctypes.data, it[1], it[2], it[3] if len(it) > 3 else 0)
        for s, o, it in zip(srcs, outs, items)
    ])
    rc = lib().mob200_decode_batch_host(ctx.handle, arr, len(items))
    if rc < 0:
        raise RuntimeError(f"mob200_decode_batch_host failed ({rc})")
    return [o[: it[1] * it[2]] for o, it in zip(outs, items)], [int(arr[i].status) for i in range(len(items))]


def filter_device(filter, device_ptr: int, count: int, stride: int, cuda_stream: int = 0) -> None:
    rc = lib().mob200_filter_device(_filter_id(filter), c_void_p(device_ptr), count, stride, c_void_p(cuda_stream))
    if rc != 0:
        raise RuntimeError(f"mob200_filter_device failed ({rc})")


def version() -> str:
    return lib().mob200_version().decode()


# ---------------------------------------------------------------------------------------------
# index streams (reference src/meshoptimizer.h:344-376; JS wrapper js/meshopt_decoder.mjs decodeIndexBuffer /
# decodeIndexSequence)
# ---------------------------------------------------------------------------------------------

def decode_index_version(source) -> int:
    """``meshopt_decodeIndexVersion``: 0 or 1, -1 for an invalid header."""
    src = _as_u8(source)
    return int(lib().meshopt_decodeIndexVersion(src.ctypes.data if src.size else None, src.size))


def _decode_index_rc(symbol: str, count: int, size: int, source):
    if size not in (2, 4):
        raise ValueError("index size must be 2 or 4")
    src = _as_u8(source)
    out = np.zeros(max(count, 1), dtype=np.uint16 if size == 2 else np.uint32)
    rc = getattr(lib(), symbol)(out.ctypes.data, count, size, src.ctypes.data if src.size else None, src.size)
    return int(rc), out[:count]


def decode_index_buffer_rc(count: int, size: int, source):
    """``meshopt_decodeIndexBuffer`` on host memory; returns (return code, indices)."""
    if count % 3:
        raise ValueError("index count of a triangle list must be a multiple of 3")
    return _decode_index_rc("meshopt_decodeIndexBuffer", count, size, source)


def decode_index_sequence_rc(count: int, size: int, source):
    """``meshopt_decodeIndexSequence`` on host memory; returns (return code, indices)."""
    return _decode_index_rc("meshopt_decodeIndexSequence", count, size, source)


def decode_index_buffer(count: int, size: int, source) -> np.ndarray:
    rc, out = decode_index_buffer_rc(count, size, source)
    if rc != 0:
        raise RuntimeError(f"Malformed buffer data: {rc}")
    return out


def decode_index_sequence(count: int, size: int, source) -> np.ndarray:
    rc, out = decode_index_sequence_rc(count, size, source)
    if rc != 0:
        raise RuntimeError(f"Malformed buffer data: {rc}")
    return out


def decode_index_batch_host(items: Sequence[tuple], ctx: Optional[Context] = None):
    """items: (source bytes, index_count, index_size, kind) -> (list of index arrays, list of return codes);
    one kernel launch for the whole batch (``mob200_decode_index_batch_host``)."""
    ctx = ctx or default_context()
    n = len(items)
    arr = (IndexStream * max(n, 1))()
    keep, outs = [], []
    for i, (src, count, size, kind) in enumerate(items):
        s = _as_u8(src)
        keep.append(s)
        o = np.zeros(max(count, 1), dtype=np.uint16 if size == 2 else np.uint32)
        outs.append(o)
        arr[i].src = s.ctypes.data if s.size else None
        arr[i].src_size = s.size
        arr[i].dst = o.ctypes.data
        arr[i].index_count = count
        arr[i].index_size = size
        arr[i].kind = kind
    rc = lib().mob200_decode_index_batch_host(ctx.handle, arr, n)
    if rc < 0:
        raise RuntimeError(f"mob200_decode_index_batch_host failed ({rc})")
    return [o[: it[1]] for o, it in zip(outs, items)], [arr[i].status for i in range(n)]


# ---------------------------------------------------------------------------------------------
# meshlets (reference src/meshoptimizer.h:349-350)
# ---------------------------------------------------------------------------------------------

def _meshlet_outputs(vertex_count, vertex_size, triangle_count, triangle_size):
    if vertex_size not in (2, 4) or triangle_size not in (3, 4) or vertex_count > 256 or triangle_count > 256:
        raise ValueError("meshlet: counts <= 256, vertex_size 2|4, triangle_size 3|4")
    v = np.zeros(max(vertex_count, 1), dtype=np.uint16 if vertex_size == 2 else np.uint32)
    t = np.zeros((max(triangle_count, 1), 3), dtype=np.uint8) if triangle_size == 3 else np.zeros(max(triangle_count, 1), dtype=np.uint32)
    return v, t


def decode_meshlet_rc(vertex_count: int, vertex_size: int, triangle_count: int, triangle_size: int, source):
    """``meshopt_decodeMeshlet`` on host memory -> (return code, vertex references, triangles (u8[n,3] or packed u32[n]))"""
    v, t = _meshlet_outputs(vertex_count, vertex_size, triangle_count, triangle_size)
    src = _as_u8(source)
    rc = lib().meshopt_decodeMeshlet(v.ctypes.data, vertex_count, vertex_size, t.ctypes.data, triangle_count, triangle_size, src.ctypes.data if src.size else None, src.size)
    return int(rc), v[:vertex_count], t[:triangle_count]


def decode_meshlet_batch_host(items: Sequence[tuple], ctx: Optional[Context] = None):
    """items: (source, vertex_count, vertex_size, triangle_count, triangle_size) -> (list of (vertices, triangles), codes);
    one kernel launch for the whole batch (``mob200_decode_meshlet_batch_host``)."""
    ctx = ctx or default_context()
    n = len(items)
    arr = (Meshlet * max(n, 1))()
    keep, outs = [], []
    for i, (src, vc, vs, tc, ts) in enumerate(items):
        s = _as_u8(src)
        v, t = _meshlet_outputs(vc, vs, tc, ts)
        keep.append(s)
        outs.append((v[:vc], t[:tc]))
        arr[i].src, arr[i].src_size = (s.ctypes.data if s.size else None), s.size
        arr[i].vertices, arr[i].vertex_count, arr[i].vertex_size = v.ctypes.data, vc, vs
        arr[i].triangles, arr[i].triangle_count, arr[i].triangle_size = t.ctypes.data, tc, ts
    rc = lib().mob200_decode_meshlet_batch_host(ctx.handle, arr, n)
    if rc < 0:
        raise RuntimeError(f"mob200_decode_meshlet_batch_host failed ({rc})")
    return outs, [arr[i].status for i in range(n)]


# ---------------------------------------------------------------------------------------------
# glTF bufferView front-end (reference gltf/parsegltf.cpp:561-627)
# ---------------------------------------------------------------------------------------------

def gltf_scan(data):
    """``mob200_gltf_scan``: (views ctypes array, buffer byteLengths, info) of a .glb / .gltf JSON blob."""
    src = _as_u8(data)
    info = GltfInfo()
    rc = lib().mob200_gltf_scan(src.ctypes.data, src.size, None, 0, None, 0, ctypes.byref(info))
    if rc != 0:
        raise ValueError(f"not a glTF asset ({rc})")
    views = (GltfView * max(1, info.view_count))()
    sizes = (c_size_t * max(1, info.buffer_count))()
    rc = lib().mob200_gltf_scan(src.ctypes.data, src.size, views, info.view_count, sizes, info.buffer_count, ctypes.byref(info))
    assert rc == 0
    return views, [int(sizes[i]) for i in range(info.buffer_count)], info


def gltf_decode_host(data, external_buffers: Optional[dict] = None, ctx: Optional[Context] = None):
    """Decompress every meshopt-compressed bufferView of a .glb in one batched device decode.
    Returns (dict buffer index -> uint8 array with the decompressed views at their byteOffset, views, info).
    buffers[0] is the BIN chunk of the .glb; other source buffers come from ``external_buffers``."""
    ctx = ctx or default_context()
    src = _as_u8(data)
    views, sizes, info = gltf_scan(src)
    n = info.view_count
    nb = max(1, info.buffer_count)
    sources = dict(external_buffers or {})
    if info.bin_size:
        sources.setdefault(0, src[info.bin_offset : info.bin_offset + info.bin_size])
    bufs = (c_void_p * nb)()
    lens = (c_size_t * nb)()
    keep = []
    for i in range(info.buffer_count):
        b = sources.get(i)
        if b is not None:
            b = _as_u8(b)
            keep.append(b)
            bufs[i] = b.ctypes.data
            lens[i] = b.size
    outputs = {}
    outs = (c_void_p * nb)()
    out_lens = (c_size_t * nb)()
    for k in range(n):
        d = views[k].dst_buffer
        if d < info.buffer_count and d not in outputs:
            outputs[d] = np.zeros(max(sizes[d], 1), dtype=np.uint8)
            outs[d] = outputs[d].ctypes.data
            out_lens[d] = sizes[d]
    rc = lib().mob200_gltf_decode_host(ctx.handle, views, n, info.buffer_count, bufs, lens, outs, out_lens)
    if rc < 0:
        raise RuntimeError(f"mob200_gltf_decode_host failed ({rc})")
    return outputs, views, info
