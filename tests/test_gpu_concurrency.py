"""-m gpu: the drop-in symbols are callable from many host threads at once (reference contract: reentrant, no locks,
SURVEY.md section 8b "Threading"; callers: js/meshopt_decoder.mjs:82-162 worker pool, gltf/parsegltf.cpp:561-627).
Every thread checks its own results bit for bit; the library serves the calls from a pool of contexts."""
import threading

import numpy as np
import pytest

from oracle import loader, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import meshoptimizer_b200 as m

    m.lib()
    return m


def test_dropin_symbols_from_sixteen_threads(mb, ref_vectors, checker):
    data = ref_vectors
    meta = data["codec_meta"]
    streams = [(data[f"codec_{i}_enc"], int(c), int(vs), data[f"codec_{i}_dec"]) for i, (c, vs, _, _) in enumerate(meta)]
    big = workloads.c2(total=1 << 17, seg=None) if loader.have_ref() else None
    filters = [("exp", 12, data["filter_exp15_in"], data["filter_exp15_out"]), ("quat", 8, data["filter_quat12_in"], data["filter_quat12_out"]),
               ("oct", 8, data["filter_oct12_in"], data["filter_oct12_out"])] if "filter_quat12_in" in data.files else [("exp", 12, data["filter_exp15_in"], data["filter_exp15_out"])]
    errors = []
    start = threading.Barrier(16)

    def worker(tid):
        try:
            rng = np.random.default_rng(tid)
            start.wait()
            for it in range(12):
                enc, count, vs, want = streams[int(rng.integers(len(streams)))]
                rc, out = mb.decode_vertex_buffer_rc(count, vs, enc)
                assert rc == 0 and np.array_equal(out, want), ("codec", tid, it)
                name, stride, fin, fout = filters[int(rng.integers(len(filters)))]
                buf = fin.copy()
                getattr(mb, "decode_filter_" + name)(buf, buf.size // stride, stride)
                assert np.array_equal(buf, fout), ("filter", name, tid, it)
                if big is not None and it % 4 == tid % 4:
                    rc, out = mb.decode_vertex_buffer_rc(int(big.counts[0]), 32, big.stream(0))
                    assert rc == 0 and np.array_equal(out, big.source), ("big", tid, it)
                # a malformed stream keeps its reference code under concurrency, too
                rc, _ = mb.decode_vertex_buffer_rc(count, vs, enc[: max(1, enc.size // 2)])
                assert rc == checker.decode_vertex_buffer(count, vs, enc[: max(1, enc.size // 2)])[0]
        except BaseException as e:  # noqa: BLE001 - reported in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not any(t.is_alive() for t in threads), "a caller is stuck"
    assert not errors, errors[:3]


def test_host_batches_from_several_threads_with_their_own_contexts(mb):
    """explicit contexts on different threads: the persistent kernels of concurrent batches are chained per device"""
    if not loader.have_ref():
        pytest.skip("needs the reference encoder")
    w = workloads.c2(total=1 << 18, seg=1 << 12)
    want = w.source
    errors = []

    def worker(tid):
        try:
            ctx = mb.Context(0)
            for _ in range(3):
                outs, rcs = mb.decode_batch_host(w.harness_streams(), ctx=ctx)
                assert all(r == 0 for r in rcs) and np.array_equal(np.concatenate(outs), want)
            ctx.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not any(t.is_alive() for t in threads), "a caller is stuck"
    assert not errors, errors[:3]


def test_multi_device_batch_entry_point(mb):
    """mob200_decode_batch_multi_host: LPT partition + one host thread and context per listed device (here the one GPU
    of the test box, listed once and twice: the second form runs two shards concurrently on the same device)"""
    import ctypes

    if not loader.have_ref():
        pytest.skip("needs the reference encoder")
    w = workloads.merge("mixed", [workloads.c2(total=1 << 17, seg=1 << 13), workloads.c3("quat12", count=60_000, seg=7_000), workloads.c4(500)])
    want = workloads.expected_outputs(w)
    port = loader.port()
    sidecars = [port.block_offsets(int(w.counts[i]), int(w.vertex_sizes[i]), w.stream(i))[1] for i in range(w.n)]
    for devices, with_side in (([0], False), ([0, 0], True)):
        outs = [np.zeros(max(1, int(w.counts[i]) * int(w.vertex_sizes[i])), np.uint8) for i in range(w.n)]
        srcs = [np.ascontiguousarray(w.stream(i)) for i in range(w.n)]
        arr = mb.make_streams([(srcs[i].ctypes.data, srcs[i].size, outs[i].ctypes.data, int(w.counts[i]), int(w.vertex_sizes[i]), int(w.filters[i])) for i in range(w.n)])
        side = None
        if with_side:
            side = (ctypes.c_void_p * w.n)()
            for i, sc in enumerate(sidecars):
                side[i] = sc.ctypes.data if sc.size else None
        devs = (ctypes.c_int * len(devices))(*devices)
        ms = (ctypes.c_float * len(devices))()
        rc = mb.lib().mob200_decode_batch_multi_host(devs, len(devices), arr, w.n, side, ms)
        assert rc == 0 and all(arr[i].status == 0 for i in range(w.n))
        for i in range(w.n):
            nb = int(w.counts[i]) * int(w.vertex_sizes[i])
            assert np.array_equal(outs[i][:nb], want[i]), (devices, i)
