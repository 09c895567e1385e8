"""ctypes front-end for the two CPU checkers.  TEST / BENCH INFRASTRUCTURE ONLY.

* ``port()``  -> oracle/_build/libmeshopt_oracle.so : our C restatement (oracle_* symbols)
* ``ref()``   -> oracle/_ref/libmeshopt_ref.so      : the unmodified reference (meshopt_* symbols),
                 present only when it was built in a container that has /root/reference

Only tests/, __graft_entry__.smoke() and bench.py (input generation with the reference encoder,
cpu_baseline leg, --impl reference arm) may import this module.  The product package
``meshoptimizer_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libmeshopt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmeshopt_ref.so")
REF_AVX_SO = os.path.join(HERE, "_ref", "libmeshopt_ref_avx.so")  # same sources, -mavx: CPU-baseline timing only

FILTER_NONE, FILTER_OCT, FILTER_QUAT, FILTER_EXP, FILTER_COLOR = 0, 1, 2, 3, 4
FILTER_NAMES = {"none": 0, "oct": 1, "quat": 2, "exp": 3, "color": 4}


class HarnessMeshlet(ctypes.Structure):
    _fields_ = [
        ("src", c_void_p), ("src_size", c_size_t),
        ("vertices", c_void_p), ("vertex_count", c_size_t), ("vertex_size", c_size_t),
        ("triangles", c_void_p), ("triangle_count", c_size_t), ("triangle_size", c_size_t),
        ("status", c_int),
    ]


class HarnessStream(ctypes.Structure):
    _fields_ = [
        ("src", c_void_p),
        ("src_size", c_size_t),
        ("dst", c_void_p),
        ("vertex_count", c_size_t),
        ("vertex_size", c_size_t),
        ("filter", c_int),
        ("status", c_int),
    ]


def build(force: bool = False) -> None:
    """Compile the port (always) and the reference library (when /root/reference exists)."""
    if force or not os.path.exists(PORT_SO) or _stale(PORT_SO):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_SO) or _stale(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_AVX_SO) or _stale(REF_AVX_SO)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref_avx"])


def _stale(so: str) -> bool:
    t = os.path.getmtime(so)
    srcs = ["vertexcodec_oracle.c", "vertexfilter_oracle.c", "indexcodec_oracle.c", "meshletcodec_oracle.c", "harness.cpp", "Makefile"]
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in srcs)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray, memoryview)) else a).view(np.uint8).reshape(-1)


class _Lib:
    """Common wrapper: decode + filters + multi-threaded harness, for either library."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.path = path
        self.kind = kind
        self.lib = ctypes.CDLL(path, mode=os.RTLD_LOCAL)  # RTLD_LOCAL: never collide with our own meshopt_* exports
        L = self.lib
        self._decode = getattr(L, prefix + "decodeVertexBuffer")
        self._decode.restype = c_int
        self._decode.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
        self._version = getattr(L, prefix + "decodeVertexVersion")
        self._version.restype = c_int
        self._version.argtypes = [c_void_p, c_size_t]
        self._filters = {}
        for name in ("Oct", "Quat", "Exp", "Color"):
            f = getattr(L, prefix + "decodeFilter" + name)
            f.restype = None
            f.argtypes = [c_void_p, c_size_t, c_size_t]
            self._filters[name.lower()] = f
        self._index = {}
        for name in ("decodeIndexBuffer", "decodeIndexSequence"):
            f = getattr(L, prefix + name)
            f.restype = c_int
            f.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
            self._index[name] = f
        self._index_version = getattr(L, prefix + "decodeIndexVersion")
        self._index_version.restype = c_int
        self._index_version.argtypes = [c_void_p, c_size_t]
        self._meshlet = getattr(L, prefix + "decodeMeshlet")
        self._meshlet.restype = c_int
        self._meshlet.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
        L.harness_decode_meshlets_mt.restype = c_double
        L.harness_decode_meshlets_mt.argtypes = [POINTER(HarnessMeshlet), c_size_t, c_int, c_int]
        L.harness_decode_mt.restype = c_double
        L.harness_decode_mt.argtypes = [POINTER(HarnessStream), c_size_t, c_int, c_int, POINTER(c_double)]
        L.harness_hw_threads.restype = c_int
        L.harness_gen_grid.argtypes = [c_void_p, c_int]
        L.harness_gen_js16.argtypes = [c_void_p, c_size_t]
        L.harness_gen_c2.argtypes = [c_void_p, c_uint64, c_size_t, c_int]

    # -- single calls ---------------------------------------------------------------------------
    def decode_vertex_buffer(self, vertex_count: int, vertex_size: int, data) -> tuple[int, np.ndarray]:
        src = _u8(data)
        out = np.zeros(max(vertex_count * vertex_size, 1), dtype=np.uint8)
        rc = self._decode(out.ctypes.data, vertex_count, vertex_size, src.ctypes.data if src.size else None, src.size)
        return rc, out[: vertex_count * vertex_size]

    def decode_vertex_version(self, data) -> int:
        src = _u8(data)
        return self._version(src.ctypes.data if src.size else None, src.size)

    def decode_index(self, kind: str, index_count: int, index_size: int, data) -> tuple[int, np.ndarray]:
        """kind: 'triangles' (decodeIndexBuffer) or 'sequence' (decodeIndexSequence)"""
        src = _u8(data)
        out = np.zeros(max(index_count * index_size, 4), dtype=np.uint8)
        f = self._index["decodeIndexBuffer" if kind == "triangles" else "decodeIndexSequence"]
        rc = f(out.ctypes.data, index_count, index_size, src.ctypes.data if src.size else None, src.size)
        return rc, out[: index_count * index_size].view(np.uint16 if index_size == 2 else np.uint32)

    def decode_meshlet(self, vertex_count: int, vertex_size: int, triangle_count: int, triangle_size: int, data):
        """-> (rc, vertex references u16|u32, triangles: u8[count, 3] or packed u32[count])"""
        src = _u8(data)
        v = np.zeros(max(vertex_count, 1) + 4, dtype=np.uint16 if vertex_size == 2 else np.uint32)
        t = np.zeros(max(triangle_count, 1) * 4 + 16, dtype=np.uint8)
        rc = self._meshlet(v.ctypes.data, vertex_count, vertex_size, t.ctypes.data, triangle_count, triangle_size, src.ctypes.data if src.size else None, src.size)
        tri = t[: triangle_count * 3].reshape(-1, 3) if triangle_size == 3 else t[: triangle_count * 4].view(np.uint32)
        return rc, v[:vertex_count], tri

    def decode_meshlets_mt(self, items, threads: int, passes: int = 1):
        """items: (src, vertex_count, vertex_size, triangle_count, triangle_size); returns best seconds, statuses"""
        n = len(items)
        arr = (HarnessMeshlet * n)()
        keep = []
        for i, (src, vc, vs, tc, ts) in enumerate(items):
            s = _u8(src)
            v = np.zeros(vc * vs + 16, np.uint8)
            t = np.zeros(tc * ts + 16, np.uint8)
            keep += [s, v, t]
            arr[i].src, arr[i].src_size = s.ctypes.data, s.size
            arr[i].vertices, arr[i].vertex_count, arr[i].vertex_size = v.ctypes.data, vc, vs
            arr[i].triangles, arr[i].triangle_count, arr[i].triangle_size = t.ctypes.data, tc, ts
        best = self.lib.harness_decode_meshlets_mt(arr, n, threads, passes)
        return best, [arr[i].status for i in range(n)]

    def decode_index_version(self, data) -> int:
        src = _u8(data)
        return self._index_version(src.ctypes.data if src.size else None, src.size)

    def decode_filter(self, name: str, buf: np.ndarray, count: int, stride: int) -> np.ndarray:
        out = np.ascontiguousarray(buf).view(np.uint8).reshape(-1).copy()
        assert out.size >= count * stride
        self._filters[name](out.ctypes.data, count, stride)
        return out

    # -- batch, multi-threaded (cpu baseline) --------------------------------------------------------
    def decode_batch_mt(self, streams, threads: int, passes: int = 1):
        """streams: list of (src u8 array, vertex_count, vertex_size, filter id).
        Returns (best_seconds, per_pass_seconds, outputs list, statuses)."""
        n = len(streams)
        arr = (HarnessStream * n)()
        outs = []
        keep = []
        for i, (src, count, vs, filt) in enumerate(streams):
            s = _u8(src)
            keep.append(s)
            o = np.empty(max(count * vs, 1), dtype=np.uint8)
            o.fill(0)  # pre-faulted: the timed passes never pay for first-touch page faults
            outs.append(o)
            arr[i].src = s.ctypes.data
            arr[i].src_size = s.size
            arr[i].dst = o.ctypes.data
            arr[i].vertex_count = count
            arr[i].vertex_size = vs
            arr[i].filter = filt
        times = (c_double * passes)()
        best = self.lib.harness_decode_mt(arr, n, threads, passes, times)
        return best, list(times), [o[: c * v] for o, (_, c, v, _) in zip(outs, streams)], [arr[i].status for i in range(n)]

    def block_offsets(self, vertex_count: int, vertex_size: int, data):
        """(return code, block-offset table uint32[nblocks + 1]) -- port only (the reference has no such entry point)"""
        f = self.lib.oracle_vertexBlockOffsets
        f.restype = c_int
        f.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_size_t]
        src = _u8(data)
        bv = min(256, (8192 // vertex_size) & ~15)
        nb = (vertex_count + bv - 1) // bv
        out = np.zeros(nb + 1 if nb else 1, dtype=np.uint32)
        rc = f(out.ctypes.data, vertex_count, vertex_size, src.ctypes.data if src.size else None, src.size)
        return int(rc), out[: nb + 1 if nb else 0]

    def hw_threads(self) -> int:
        return int(self.lib.harness_hw_threads())

    # -- synthetic vertex data -------------------------------------------------------------------------
    def gen_grid(self, side: int) -> np.ndarray:
        out = np.empty(((side + 1) * (side + 1), 16), dtype=np.uint16)
        self.lib.harness_gen_grid(out.ctypes.data, side)
        return out

    def gen_js16(self, vertex_count: int) -> np.ndarray:
        out = np.empty((vertex_count, 16), dtype=np.uint8)
        self.lib.harness_gen_js16(out.ctypes.data, vertex_count)
        return out

    def gen_c2(self, first: int, count: int, threads: int = 0) -> np.ndarray:
        out = np.empty((count, 16), dtype=np.uint16)
        self.lib.harness_gen_c2(out.ctypes.data, first, count, threads or self.hw_threads())
        return out


class _Ref(_Lib):
    """The unmodified reference: adds the encoders (input generation)."""

    def __init__(self, path: str = REF_SO, kind: str = "reference"):
        super().__init__(path, "meshopt_", kind)
        L = self.lib
        L.meshopt_encodeVertexBufferLevel.restype = c_size_t
        L.meshopt_encodeVertexBufferLevel.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_int]
        L.meshopt_encodeVertexBufferBound.restype = c_size_t
        L.meshopt_encodeVertexBufferBound.argtypes = [c_size_t, c_size_t]
        L.harness_encode_segments.restype = None
        L.harness_encode_segments.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
        L.harness_grid_reorder.argtypes = [c_void_p, c_int]
        for name, extra in (("Oct", []), ("Quat", []), ("Color", []), ("Exp", [c_int])):
            f = getattr(L, "meshopt_encodeFilter" + name)
            f.restype = None
            f.argtypes = [c_void_p, c_size_t, c_size_t, c_int, c_void_p] + extra

    def encode_index(self, kind: str, indices, vertex_count: int, version: int = 1) -> np.ndarray:
        """reference index encoders (input generation): kind 'triangles' or 'sequence'"""
        L = self.lib
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        L.meshopt_encodeIndexVersion.argtypes = [c_int]
        L.meshopt_encodeIndexVersion(version)
        name = "meshopt_encodeIndexBuffer" if kind == "triangles" else "meshopt_encodeIndexSequence"
        bound = getattr(L, name + "Bound")
        bound.restype = c_size_t
        bound.argtypes = [c_size_t, c_size_t]
        enc = getattr(L, name)
        enc.restype = c_size_t
        enc.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t]
        buf = np.empty(int(bound(idx.size, vertex_count)), dtype=np.uint8)
        n = enc(buf.ctypes.data, buf.size, idx.ctypes.data if idx.size else None, idx.size)
        L.meshopt_encodeIndexVersion(1)
        assert n > 0
        return buf[:n].copy()

    def encode_meshlet(self, vertices, triangles) -> np.ndarray:
        """reference meshopt_encodeMeshlet: vertices u32[<=256], triangles u8[n, 3] of local indices"""
        L = self.lib
        v = np.ascontiguousarray(vertices, dtype=np.uint32)
        t = np.ascontiguousarray(triangles, dtype=np.uint8).reshape(-1, 3)
        L.meshopt_encodeMeshletBound.restype = c_size_t
        L.meshopt_encodeMeshletBound.argtypes = [c_size_t, c_size_t]
        L.meshopt_encodeMeshlet.restype = c_size_t
        L.meshopt_encodeMeshlet.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t]
        buf = np.empty(int(L.meshopt_encodeMeshletBound(256, 256)), dtype=np.uint8)
        n = L.meshopt_encodeMeshlet(buf.ctypes.data, buf.size, v.ctypes.data if v.size else None, v.size, t.ctypes.data if t.size else None, t.shape[0])
        assert n > 0
        return buf[:n].copy()

    def encode_bound(self, vertex_count: int, vertex_size: int) -> int:
        return int(self.lib.meshopt_encodeVertexBufferBound(vertex_count, vertex_size))

    def encode_vertex_buffer(self, vertices: np.ndarray, vertex_count: int, vertex_size: int, level: int = 2, version: int = 1) -> np.ndarray:
        v = _u8(vertices)
        assert v.size >= vertex_count * vertex_size
        buf = np.empty(self.encode_bound(vertex_count, vertex_size), dtype=np.uint8)
        n = self.lib.meshopt_encodeVertexBufferLevel(buf.ctypes.data, buf.size, v.ctypes.data if v.size else None, vertex_count, vertex_size, level, version)
        assert n > 0
        return buf[:n].copy()

    def encode_segments(self, vertices: np.ndarray, vertex_size: int, firsts, counts, level: int = 2, version: int = 1, threads: int = 0):
        """Encode independent segments of one vertex array in parallel.
        Returns (blob u8, offsets u64[n], sizes u64[n]); segment i is blob[offsets[i]:offsets[i]+sizes[i]].
        Every segment starts on a 16-byte boundary of the blob."""
        v = _u8(vertices)
        firsts = np.ascontiguousarray(firsts, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        n = firsts.size
        caps = np.array([(self.encode_bound(int(c), vertex_size) + 15) & ~15 for c in counts], dtype=np.uint64)
        offs = np.zeros(n, dtype=np.uint64)
        np.cumsum(caps[:-1], out=offs[1:])
        blob = np.empty(int(caps.sum()), dtype=np.uint8)
        sizes = np.zeros(n, dtype=np.uint64)
        self.lib.harness_encode_segments(v.ctypes.data, vertex_size, firsts.ctypes.data, counts.ctypes.data, n, level, version,
                                         blob.ctypes.data, offs.ctypes.data, caps.ctypes.data, sizes.ctypes.data, threads or self.hw_threads())
        assert (sizes > 0).all()
        # compact: keep 16-byte alignment of every segment start
        new_offs = np.zeros(n, dtype=np.uint64)
        padded = (sizes + np.uint64(15)) & ~np.uint64(15)
        np.cumsum(padded[:-1], out=new_offs[1:])
        out = np.zeros(int(padded.sum()) + 16, dtype=np.uint8)
        for i in range(n):
            o, s, no = int(offs[i]), int(sizes[i]), int(new_offs[i])
            out[no : no + s] = blob[o : o + s]
        return out, new_offs, sizes

    def grid_reorder(self, vertices: np.ndarray, side: int) -> np.ndarray:
        v = np.ascontiguousarray(vertices).copy()
        self.lib.harness_grid_reorder(v.ctypes.data, side)
        return v

    def encode_filter(self, name: str, data: np.ndarray, count: int, stride: int, bits: int, mode: int = 0) -> np.ndarray:
        src = np.ascontiguousarray(data, dtype=np.float32)
        out = np.zeros(count * stride, dtype=np.uint8)
        f = getattr(self.lib, "meshopt_encodeFilter" + name.capitalize())
        if name == "exp":
            f(out.ctypes.data, count, stride, bits, src.ctypes.data, mode)
        else:
            f(out.ctypes.data, count, stride, bits, src.ctypes.data)
        return out


_port = None
_ref = None


def port() -> _Lib:
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        _port = _Lib(PORT_SO, "oracle_", "port")
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_ref_avx = None


def have_ref_avx() -> bool:
    return os.path.exists(REF_AVX_SO)


def ref_avx() -> _Ref:
    """the reference compiled with -mavx (timing only; parity always uses ref())"""
    global _ref_avx
    if _ref_avx is None:
        _ref_avx = _Ref(REF_AVX_SO, "reference")
    return _ref_avx


def ref() -> _Ref:
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libmeshopt_ref.so not built (needs /root/reference at build time)")
        _ref = _Ref()
    return _ref
