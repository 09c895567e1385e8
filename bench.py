#!/usr/bin/env python
"""bench.py -- decoded vertex GB/s of the B200 vertex-buffer decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--verts V] [--headline NAME] [--no-configs]

Headline workload (N=1): BASELINE.json configs[1] -- the v1 codec at encode level 2 on a 64 Mi-vertex, 32-byte-per-vertex
data set (2.147 GB decoded), encoded by the UNMODIFIED reference encoder (oracle/_ref) as SURVEY.md Appendix D "C2b":
1024 independently encoded streams of 65 536 vertices, each with its block-offset SIDECAR (4 bytes per block; SURVEY.md
section 8b "optional d_block_offsets sidecar", include/meshopt_b200.h section 2b).  The sidecar is INPUT, produced at
encode time (here: by the CPU checker's walk at workload generation, never by a GPU run); every step walks and verifies
every block (one walker lane per block) and decodes it: nothing is cached between steps.  The same streams WITHOUT the
sidecar (serial walk per stream), the 4096-vertex segmentation, the monolithic stream and the codecbench grid (C1a) are
measured in the same run and reported in "configs", each with its time, decoded GB/s, roofline fraction and a
full-output parity check on the device.

  value     decoded GB/s, inputs and outputs resident in HBM, CUDA events on the launching stream, max over ranks,
            K steps back to back after W warm-up steps (working set >> L2)
  e2e       same metric through the reference-facing C ABI with HOST buffers (mob200_decode_batch_host_sidecar):
            pinned host memory -> device -> decode -> pinned host memory
  roofline  dominant kernel: algorithmic bytes (encoded + sidecar read once, decoded written once) per launch / its
            CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference
            the reference's own SIMD decoder (oracle/_ref, meshopt_decodeVertexBuffer per stream; default and -mavx
            builds, the faster is reported) on all host threads of the box, same streams, outputs pre-faulted

Multi-GPU (torchrun, one rank per GPU): the job is N x 64 Mi vertices = N x 1024 independent streams, partitioned over the
ranks by the library's longest-processing-time rule (meshoptimizer_b200/sharding.py); every rank decodes its own shard;
no collective is on the data path (NCCL only provides the barrier and the max-over-ranks of the timing).  scaling = weak.
(One process driving all GPUs of a box: mob200_decode_batch_multi_host, measured by tools/bench_multi.py.)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "decoded_vertex_GBps"
UNIT = "GB/s"
HEADLINES = ("c2b_sidecar", "c2b", "seg4096", "seg4096_sidecar")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--verts", type=int, default=1 << 26, help="vertices per GPU (default 64 Mi)")
    ap.add_argument("--headline", default="c2b_sidecar", choices=HEADLINES,
                    help="c2b*: 65536-vertex streams (SURVEY Appendix D C2b); seg4096*: 4096-vertex streams; *_sidecar: block mode")
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--version", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-passes", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configurations (they only run at N=1)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def numa_bind(local: int):
    """Best effort: run this rank's host threads on the NUMA node its GPU hangs off, BEFORE any pinned memory is
    allocated (pinned pages are placed by first touch), so that the host side of the e2e arm does not cross sockets.
    Returns a short description for the JSON line."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return f"gpu {local} ({bus}): no NUMA affinity reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {local} ({bus}): node {node} has no allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"gpu {local} ({bus}) -> NUMA node {node}, {len(cpus)} CPUs"
    except Exception as e:  # noqa: BLE001 - diagnostics only
        return f"not bound ({type(e).__name__})"


def headline_shape(name: str):
    """(vertices per stream, block mode)"""
    return {"c2b_sidecar": (1 << 16, True), "c2b": (1 << 16, False), "seg4096": (1 << 12, False), "seg4096_sidecar": (1 << 12, True)}[name]


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------

def gen_vertices(first_vertex: int, verts: int, threads: int) -> np.ndarray:
    """C2 vertex data: uint8[verts * 32]"""
    from oracle import loader

    return loader.port().gen_c2(first_vertex, verts, threads).view(np.uint8).reshape(-1)


def encode_workload(v: np.ndarray, vs: int, segment, level: int, version: int, threads: int, sidecar: bool):
    """Encode v (uint8, vs bytes per vertex) with the reference encoder as independent streams of `segment` vertices
    (None: one stream).  Streams are packed back to back on 16-byte boundaries.  With sidecar=True every stream also
    gets its block-offset table from the CPU checker's walk (oracle port; what an encoder would emit)."""
    from oracle import loader

    R, P = loader.ref(), loader.port()
    verts = v.size // vs
    seg = segment or verts
    firsts = np.arange(0, verts, seg, dtype=np.uint64)
    counts = np.minimum(np.uint64(seg), np.uint64(verts) - firsts).astype(np.uint64)
    blob, offs, sizes = R.encode_segments(v, vs, firsts, counts, level, version, threads)
    n = firsts.size
    packed_off = np.zeros(n, np.uint64)
    np.cumsum(((sizes + np.uint64(15)) & ~np.uint64(15))[:-1], out=packed_off[1:])
    total = int(packed_off[-1] + ((sizes[-1] + np.uint64(15)) & ~np.uint64(15)))
    packed = np.zeros(total + 64, np.uint8)
    for i in range(n):
        o, s, p = int(offs[i]), int(sizes[i]), int(packed_off[i])
        packed[p : p + s] = blob[o : o + s]
    del blob
    bv = min(256, (8192 // vs) & ~15)
    wl = dict(blob=packed, offsets=packed_off, sizes=sizes, counts=counts, vs=vs, segment=segment, level=level, version=version,
              decoded_bytes=int(counts.sum()) * vs, encoded_bytes=int(sizes.sum()), sidecars=None,
              sidecar_bytes=int(((counts + np.uint64(bv - 1)) // np.uint64(bv) + np.uint64(1)).sum()) * 4)  # nblocks + 1 entries per stream
    if sidecar:
        def one(i):
            o, s = int(packed_off[i]), int(sizes[i])
            rc, off = P.block_offsets(int(counts[i]), vs, packed[o : o + s])
            assert rc == 0
            return off

        with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
            wl["sidecars"] = list(ex.map(one, range(n)))
        assert wl["sidecar_bytes"] == int(sum(s.size for s in wl["sidecars"])) * 4
    return wl


def build_workload(verts: int, segment: int, level: int, version: int, first_vertex: int, threads: int):
    """(diagnostic tools) C2 vertices encoded as streams of `segment` vertices, with a few original streams to check"""
    v = gen_vertices(first_vertex, verts, threads)
    wl = encode_workload(v, 32, segment, level, version, threads, False)
    n = len(wl["offsets"])
    wl["check"] = {}
    for j in (0, n // 2, n - 1):
        a = j * segment * 32
        wl["check"][j] = v[a : a + int(wl["counts"][j]) * 32].copy()
    return wl


def describe(wl, verts, block: bool):
    n = len(wl["offsets"])
    seg = f"{n} independent streams of {wl['segment']} vertices" if wl["segment"] else "ONE monolithic stream"
    side = (f" + block-offset sidecar ({wl['sidecar_bytes']} B, input; every block walked and verified by its own lane each step)" if block
            else ", serial walk per stream (no sidecar)")
    return (f"v{wl['version']} codec, encode level {wl['level']}, {verts} vertices x {wl['vs']} bytes ({wl['decoded_bytes']/1e9:.3f} GB decoded, "
            f"{wl['encoded_bytes']/1e9:.3f} GB encoded), reference-encoded as {seg}{side}")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs"""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val == "Active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (reference decoder on the host cores)
# ------------------------------------------------------------------------------------------------

def cpu_libs():
    """the reference builds that may be timed: (label, lib) -- default SSE build and the -mavx build"""
    from oracle import loader

    libs = []
    if loader.have_ref():
        libs.append(("reference (g++ -O3, default x86-64 SIMD path)", loader.ref()))
    if loader.have_ref_avx():
        libs.append(("reference (g++ -O3 -mavx, reference Makefile release-avx)", loader.ref_avx()))
    if not libs:
        libs.append(("port (oracle C restatement)", loader.port()))
    return libs


def cpu_decode(lib, wl, passes: int, threads: int, max_streams=None):
    n = len(wl["offsets"]) if max_streams is None else min(max_streams, len(wl["offsets"]))
    streams = []
    for i in range(n):
        o, s = int(wl["offsets"][i]), int(wl["sizes"][i])
        streams.append((wl["blob"][o : o + s], int(wl["counts"][i]), wl["vs"], 0))
    best, times, outs, status = lib.decode_batch_mt(streams, threads, passes)
    assert all(s == 0 for s in status)
    decoded = sum(int(wl["counts"][i]) * wl["vs"] for i in range(n))
    return dict(seconds=best, times=times, decoded_bytes=decoded, kind=lib.kind, threads=threads, n_streams=n)


def run_reference_arm(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    libs = cpu_libs()
    threads = libs[0][1].hw_threads()
    segment, block = headline_shape(args.headline)
    v = gen_vertices(0, args.verts, threads)
    wl = encode_workload(v, 32, segment, args.level, args.version, threads, False)
    del v
    # the faster build is the arm (one untimed pass each decides; outputs are pre-faulted by the harness)
    probe = [(cpu_decode(lib, wl, 1, threads)["seconds"], label, lib) for label, lib in libs]
    probe.sort(key=lambda t: t[0])
    label, lib = probe[0][1], probe[0][2]
    cpu_decode(lib, wl, max(1, args.warmup), threads)
    r = cpu_decode(lib, wl, args.steps, threads)
    per_step = float(np.mean(r["times"]))
    value = wl["decoded_bytes"] / per_step / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, wl, block),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": r["kind"], "build": label,
                         "builds_probed": {lb: wl["decoded_bytes"] / s / 1e9 for s, lb, _ in probe},
                         "sample": f"full workload, {r['n_streams']} streams, one meshopt_decodeVertexBuffer call per stream, {threads} host threads, outputs pre-faulted, mean of {args.steps} passes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, wl, block):
    n = len(wl["offsets"])
    return {
        "workload": "BASELINE configs[1] as SURVEY Appendix D " + ("C2b" if wl["segment"] == 1 << 16 else "C2 segmentation") + ": " + describe(wl, args.verts, block) + " per GPU",
        "vertex_size": 32, "vertices_per_gpu": args.verts, "streams_per_gpu": n, "segment_vertices": wl["segment"], "block_offset_sidecar": bool(block),
        "codec_version": args.version, "encode_level": args.level,
        "l2_policy": "no flush: encoded+decoded working set per step (>= 3 GB) is far larger than the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------

def source_hash() -> str:
    """hash of the kernel sources: a committed ncu DRAM-traffic figure is only quoted for the build it was captured on"""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "meshoptimizer_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()[:16]


class DeviceWorkload:
    """one encoded workload resident on the device with its plan"""

    def __init__(self, mb, ctx, dev, wl, out, block: bool):
        import torch

        self.wl, self.block = wl, block
        self.blob = torch.from_numpy(wl["blob"]).to(dev)
        n = len(wl["offsets"])
        out_lens = (wl["counts"] * np.uint64(wl["vs"]) + np.uint64(15)) & ~np.uint64(15)
        self.out_offs = np.zeros(n, np.uint64)
        np.cumsum(out_lens[:-1], out=self.out_offs[1:])
        self.out_bytes = int(out_lens.sum())
        assert self.out_bytes <= out.numel()
        self.contiguous = bool((out_lens == wl["counts"] * np.uint64(wl["vs"])).all())
        items = [(self.blob.data_ptr() + int(wl["offsets"][i]), int(wl["sizes"][i]), out.data_ptr() + int(self.out_offs[i]), int(wl["counts"][i]), wl["vs"], 0) for i in range(n)]
        torch.cuda.synchronize(dev)  # (plan.create_ms is the plan's own host time: cudaMalloc would wait for the upload above)
        self.plan = mb.Plan(ctx, mb.make_streams(items), sidecars=wl["sidecars"] if block else None)
        self.out = out

    def run(self, stream):
        self.plan.run(stream, block_parallel=self.block)

    def parity(self, expected, stream) -> bool:
        """zero the output, decode once, compare EVERY byte with the original vertices on the device"""
        import torch

        self.out[: self.out_bytes].zero_()
        self.run(stream)
        status = self.plan.status(stream)
        if not (status == 0).all():
            return False
        if self.contiguous:
            total = self.wl["decoded_bytes"]
            ok = True
            for lo in range(0, total, 1 << 28):
                hi = min(total, lo + (1 << 28))
                ok = ok and bool(torch.equal(self.out[lo:hi], expected[lo:hi]))
            return ok
        pos, ok = 0, True
        for i in range(len(self.out_offs)):
            nb = int(self.wl["counts"][i]) * self.wl["vs"]
            o = int(self.out_offs[i])
            ok = ok and bool(torch.equal(self.out[o : o + nb], expected[pos : pos + nb]))
            pos += nb
        return ok


def time_steps(dw, stream, steps, warmup, barrier):
    import torch

    for _ in range(warmup):
        dw.run(stream)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        dw.run(stream)
    ev1.record()
    barrier()
    hist = dw.plan.timing_history(min(64, steps))
    return ev0.elapsed_time(ev1) / steps, float(np.mean(hist)), float(np.min(hist))


def config_record(name, dw, verts, ms_step, kernel_mean, kernel_best, parity, peak, steps):
    wl = dw.wl
    alg = wl["encoded_bytes"] + wl["decoded_bytes"] + (wl["sidecar_bytes"] if dw.block else 0)
    return {"name": name, "workload": describe(wl, verts, dw.block), "streams": len(wl["offsets"]), "mode": "block (sidecar)" if dw.block else "serial walk",
            "steps": steps, "ms_per_step": ms_step, "kernel_ms_mean": kernel_mean, "kernel_ms_best": kernel_best,
            "decoded_GBps": wl["decoded_bytes"] / (ms_step * 1e-3) / 1e9, "algorithmic_bytes": alg,
            "roofline_frac": alg / (kernel_mean * 1e-3) / 1e9 / peak, "plan_create_ms": dw.plan.create_ms, "parity_all_bytes": bool(parity)}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import meshoptimizer_b200 as mb
    from oracle import loader

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = numa_bind(local) if world > 1 else "single GPU: not bound"
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend="nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    mb.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    host_threads = max(1, min(len(os.sched_getaffinity(0)), (os.cpu_count() or 1) // max(1, world)))
    segment, block = headline_shape(args.headline)
    t_gen = time.time()
    if world == 1:
        v = gen_vertices(0, args.verts, host_threads)
    else:
        # the job is world x verts vertices = world x (verts / segment) independent streams; every rank computes the same
        # longest-processing-time partition (meshoptimizer_b200/sharding.py: no communication) and builds only its own streams
        from meshoptimizer_b200.sharding import shard_streams, stream_cost

        per_rank = args.verts // segment
        costs = [stream_cost(int(segment * 32 * 0.43), segment, 32)] * (world * per_rank)
        mine = shard_streams(costs, world)[rank]
        assert len(mine) == per_rank
        v = np.concatenate([gen_vertices(sid * segment, segment, host_threads) for sid in mine])
    wl = encode_workload(v, 32, segment, args.level, args.version, host_threads, block)
    t_gen = time.time() - t_gen
    n = len(wl["offsets"])

    ctx = mb.Context(local)
    stream = torch.cuda.current_stream().cuda_stream
    expected = torch.from_numpy(v).to(dev)  # the original vertices: every decoded byte is compared with them on the device
    out = torch.empty(v.size + 64, dtype=torch.uint8, device=dev)

    # ---- headline: device-resident ---------------------------------------------------------------------
    dw = DeviceWorkload(mb, ctx, dev, wl, out, block)
    assert dw.parity(expected, stream), "headline workload decoded incorrectly"
    plan_launches = dw.plan.launches  # kernels per step of the mode that just ran (1: fused walk + decode; 2: team walk + block-mode decode)
    sampler = ClockSampler(local)
    sampler.start()
    ms_per_step, kernel_ms, kernel_best = time_steps(dw, stream, args.steps, max(3, args.warmup), barrier)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
    value = world * wl["decoded_bytes"] / (ms_per_step * 1e-3) / 1e9
    parity_after = dw.parity(expected, stream)
    headline_rec = config_record(args.headline, dw, args.verts, ms_per_step, kernel_ms, kernel_best, parity_after, peak, args.steps)

    alg_bytes = wl["encoded_bytes"] + wl["decoded_bytes"] + (wl["sidecar_bytes"] if block else 0)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes of one launch: from the committed ncu --set full capture of THIS build and config (profiles/dram_traffic.json);
    # a capture of another build of the kernels is not quoted
    traffic, traffic_src, traffic_stale = None, None, None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        sh = source_hash()
        for rec in json.load(open(tpath)):
            if rec.get("config") == args.headline and rec.get("verts") == args.verts and rec.get("level") == args.level and rec.get("version") == args.version:
                if rec.get("source_hash") == sh:
                    traffic, traffic_src, traffic_stale = rec["dram_bytes_per_launch"], rec.get("source"), False
                elif traffic is None:
                    traffic_src, traffic_stale = f"{rec.get('source')} was captured on kernel sources {rec.get('source_hash')}, this build is {sh}: not quoted", True
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src, "traffic_stale": traffic_stale,
                "kernel": "decode_kernel (one fused persistent kernel per step: walker, producer and decoder warps)", "kernel_ms": kernel_ms, "kernel_ms_best": kernel_best,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}

    # ---- secondary configurations (N=1 only): the segment-size curve and the configs as literally written -------------
    configs = [headline_rec]
    if world == 1 and not args.no_configs:
        def measure(name, wl2, block2, steps, warmup, exp=expected, verts=args.verts, outbuf=out):
            d2 = DeviceWorkload(mb, ctx, dev, wl2, outbuf, block2)
            ok = d2.parity(exp, stream)
            ms, kmean, kbest = time_steps(d2, stream, steps, warmup, barrier)
            configs.append(config_record(name, d2, verts, ms, kmean, kbest, ok, peak, steps))
            del d2
            torch.cuda.empty_cache()

        other_seg = 1 << 12 if segment == 1 << 16 else 1 << 16
        # the headline streams in the other mode
        wl_alt = wl if block else encode_workload(v, 32, segment, args.level, args.version, host_threads, True)
        measure(("c2b" if segment == 1 << 16 else "seg4096") + ("" if block else "_sidecar"), wl_alt, not block, 5, 3)
        del wl_alt
        wl2 = encode_workload(v, 32, other_seg, args.level, args.version, host_threads, True)
        nm = "c2b" if other_seg == 1 << 16 else "seg4096"
        measure(nm, wl2, False, 10, 3)
        measure(nm + "_sidecar", wl2, True, 10, 3)
        del wl2
        # C2a: one monolithic stream.  With its sidecar at full size; the first-time serial walk on the first 16 Mi vertices
        # (a single chain: seconds at full size)
        wl3 = encode_workload(v, 32, None, args.level, args.version, host_threads, True)
        measure("c2a_monolithic_sidecar", wl3, True, 5, 2)
        del wl3
        mono_verts = min(args.verts, 1 << 24)
        wl4 = encode_workload(v[: mono_verts * 32], 32, None, args.level, args.version, host_threads, False)
        measure("c2a_monolithic_16Mi_first_decode", wl4, False, 1, 0, verts=mono_verts)
        del wl4
        # C1a: the codecbench grid, 1001^2 x 32 bytes, one stream, v0 (BASELINE configs[0]) and v1
        R = loader.ref()
        grid = R.grid_reorder(R.gen_grid(1000), 1000).view(np.uint8).reshape(-1)
        gexp = torch.from_numpy(grid).to(dev)
        gout = torch.empty(grid.size + 64, dtype=torch.uint8, device=dev)
        for ver, lvl in ((0, 0), (1, 2)):
            wg = encode_workload(grid, 32, None, lvl, ver, host_threads, True)
            measure(f"c1a_grid_v{ver}_first_decode", wg, False, 2, 1, exp=gexp, verts=grid.size // 32, outbuf=gout)
            measure(f"c1a_grid_v{ver}_sidecar", wg, True, 10, 3, exp=gexp, verts=grid.size // 32, outbuf=gout)
        del gexp, gout

    # ---- C3: gltfpack-style filtered streams (BASELINE configs[2]), decode + FUSED filter, SURVEY Appendix D shapes -------
    c3_table = None
    if world == 1 and not args.no_configs:
        from oracle import workloads

        c3_table = []
        P = loader.port()
        for kind in workloads.C3_KINDS:
            w3 = workloads.c3(kind, count=1 << 24, seg=1 << 16, version=1, level=2)
            want = torch.from_numpy(np.concatenate(workloads.expected_outputs(w3))).to(dev)
            vs3 = int(w3.vertex_sizes[0])
            sc3 = [P.block_offsets(int(w3.counts[i]), vs3, w3.stream(i))[1] for i in range(w3.n)]
            wl3 = dict(blob=w3.blob, offsets=w3.offsets, sizes=w3.sizes, counts=w3.counts, vs=vs3, segment=1 << 16, level=2, version=1,
                       decoded_bytes=w3.decoded_bytes, encoded_bytes=w3.encoded_bytes, sidecars=sc3, sidecar_bytes=int(sum(x.size for x in sc3)) * 4)
            o3 = torch.zeros(w3.decoded_bytes + 64, dtype=torch.uint8, device=dev)
            for blk in (False, True):
                n3 = w3.n
                offs3 = np.zeros(n3, np.uint64)
                np.cumsum((w3.counts * np.uint64(vs3))[:-1], out=offs3[1:])
                blob3 = torch.from_numpy(w3.blob).to(dev)
                items3 = [(blob3.data_ptr() + int(w3.offsets[i]), int(w3.sizes[i]), o3.data_ptr() + int(offs3[i]), int(w3.counts[i]), vs3, int(w3.filters[i])) for i in range(n3)]
                plan3 = mb.Plan(ctx, mb.make_streams(items3), sidecars=sc3 if blk else None)
                o3.zero_()
                plan3.run(stream, block_parallel=blk)
                st3 = plan3.status(stream)
                got = o3[: w3.decoded_bytes]
                if vs3 == 4 and w3.meta["filter_name"] in ("oct", "color"):  # the two lanes with a stated <= 1 LSB tolerance
                    dd = (got.to(torch.int16) - want.to(torch.int16)).abs()
                    ok3 = bool((torch.minimum(dd, 256 - dd) <= 1).all())
                else:
                    ok3 = bool(torch.equal(got, want))
                for _ in range(3):
                    plan3.run(stream, block_parallel=blk)
                for _ in range(8):
                    plan3.run(stream, block_parallel=blk)
                torch.cuda.synchronize()
                kms = float(np.min(plan3.timing_history(8)))
                alg3 = w3.encoded_bytes + w3.decoded_bytes + (wl3["sidecar_bytes"] if blk else 0)
                c3_table.append({"kind": kind, "filter": w3.meta["filter_name"], "vertex_size": vs3, "streams": n3, "elements": int(w3.counts.sum()),
                                 "mode": "block (sidecar)" if blk else "serial walk", "kernel_ms_best": kms, "decoded_GBps": w3.decoded_bytes / kms / 1e6,
                                 "traffic_GBps": alg3 / kms / 1e6, "roofline_frac": alg3 / kms / 1e6 / peak, "parity": bool(ok3 and (st3 == 0).all()),
                                 "tolerance": "<= 1 LSB" if vs3 == 4 and w3.meta["filter_name"] in ("oct", "color") else "bit-exact"})
                del plan3, blob3
            del o3, want
            torch.cuda.empty_cache()

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        import ctypes

        del dw
        h_in = torch.from_numpy(wl["blob"]).pin_memory()
        h_out = torch.empty(v.size + 64, dtype=torch.uint8).pin_memory()
        out_lens = (wl["counts"] * np.uint64(32) + np.uint64(15)) & ~np.uint64(15)
        out_offs = np.zeros(n, np.uint64)
        np.cumsum(out_lens[:-1], out=out_offs[1:])
        hitems = [(h_in.data_ptr() + int(wl["offsets"][i]), int(wl["sizes"][i]), h_out.data_ptr() + int(out_offs[i]), int(wl["counts"][i]), 32, 0) for i in range(n)]
        harr = mb.make_streams(hitems)
        side = None
        if block:
            side = (ctypes.c_void_p * n)()
            for i, sc in enumerate(wl["sidecars"]):
                side[i] = sc.ctypes.data
        del out
        torch.cuda.empty_cache()
        call = lambda: mb.lib().mob200_decode_batch_host_sidecar(ctx.handle, harr, n, side)
        rc = call()  # warm-up (allocates device arenas)
        assert rc == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rc = call()
            assert rc == 0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_s = (t1 - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_ok = bool(np.array_equal(h_out.numpy()[: v.size], v)) if bool((out_lens == wl["counts"] * np.uint64(32)).all()) else None
        assert e2e_ok is not False, "e2e: decoded bytes differ from the original vertices"
        e2e = {"value": world * wl["decoded_bytes"] / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": wl["encoded_bytes"] + (wl["sidecar_bytes"] if block else 0),
               "d2h_bytes_per_step": wl["decoded_bytes"], "ms_per_step": e2e_s * 1e3, "parity_all_bytes": e2e_ok,
               "api": "mob200_decode_batch_host_sidecar (pinned host buffers)"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        libs = cpu_libs()
        threads = libs[0][1].hw_threads()
        runs = []
        for label, lib in libs:
            r = cpu_decode(lib, wl, args.cpu_passes, threads)
            runs.append((r["decoded_bytes"] / r["seconds"] / 1e9, label, lib, r))
        runs.sort(key=lambda t: -t[0])
        gbps, label, lib, r = runs[0]
        cpu = {"value": gbps, "unit": UNIT, "cores": threads, "kind": r["kind"], "build": label, "builds": {lb: g for g, lb, _, _ in runs},
               "sample": f"full workload ({r['n_streams']} streams, {r['decoded_bytes']/1e9:.2f} GB decoded), best of {args.cpu_passes} passes, {threads} host threads, outputs pre-faulted"}
        r1 = cpu_decode(lib, wl, 1, 1, max_streams=max(1, (1 << 23) // (segment or (1 << 23))))
        cpu["single_thread_value"] = r1["decoded_bytes"] / r1["seconds"] / 1e9

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, wl, block), "clocks": clocks, "e2e": e2e, "gpu_launches": int(args.steps * plan_launches),
            "roofline": roofline, "cpu_baseline": cpu, "configs": configs, "c3_fused_filters": c3_table,
            "notes": {"generation_seconds": t_gen, "numa": numa, "plan_create_ms": headline_rec["plan_create_ms"], "parity_all_bytes": headline_rec["parity_all_bytes"],
                      "kernels_per_step": ["decode_kernel (one persistent kernel: walker, producer and decoder warps)"] if plan_launches == 1 else
                                          ["walk_team_kernel (offsets-only walk, one CTA per stream)", "decode_kernel (block mode)"],
                      "configs_key": "every entry: one fused kernel launch per step over the whole workload, device-resident, CUDA events; roofline_frac = algorithmic bytes / mean kernel time / peak; parity_all_bytes = every decoded byte compared on the device with the original vertices"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
