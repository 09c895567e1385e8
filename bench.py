#!/usr/bin/env python
"""bench.py -- decoded vertex GB/s of the B200 vertex-buffer decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--verts V] [--segment S]

Workload (N=1): BASELINE.json configs[1] -- the v1 codec at encode level 2 on a 64 Mi-vertex,
32-byte-per-vertex data set (2.147 GB decoded), produced by the UNMODIFIED reference encoder
(oracle/_ref) as independently encoded buffer ranges of `--segment` vertices each (SURVEY.md section
7.3 H1: a stream carries no block index, so the parallel unit is the independently encoded stream).
A "step" is one pass of the hot path (one persistent kernel: walker, producer and decoder warps) over the whole batch.

  value     decoded GB/s, inputs and outputs resident in HBM, CUDA events on the launching stream,
            max over ranks, K steps back to back after W warm-up steps (working set >> L2)
  e2e       same metric through the reference-facing C ABI with HOST buffers
            (mob200_decode_batch_host): pinned host memory -> device -> decode -> pinned host memory
  roofline  dominant kernel: algorithmic bytes (encoded read once + decoded written once) per launch
            / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference
            the reference's own SIMD decoder (oracle/_ref, meshopt_decodeVertexBuffer per stream) on
            all host threads of the box, same streams

Multi-GPU (torchrun, one rank per GPU): every rank decodes its own shard of independent streams
(different vertices per rank); no collective is on the data path (NCCL only provides the barrier and
the max-over-ranks of the timing).  scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "decoded_vertex_GBps"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--verts", type=int, default=1 << 26, help="vertices per GPU (default 64 Mi)")
    ap.add_argument("--segment", type=int, default=1 << 12, help="vertices per independently encoded stream")
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--version", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-passes", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# workload: generated in chunks so that host memory stays bounded (vertices are not kept)
# ------------------------------------------------------------------------------------------------

def build_workload(verts: int, segment: int, level: int, version: int, first_vertex: int, threads: int):
    """Returns dict(blob, offsets, sizes, counts, decoded_bytes, encoded_bytes, check) where `check`
    holds the original bytes of a few streams for a post-run parity check."""
    from oracle import loader

    R, P = loader.ref(), loader.port()
    chunk = max(segment, (1 << 22) // segment * segment)  # ~4 Mi vertices (128 MB) per chunk
    blobs, offs, sizes, counts = [], [], [], []
    check = {}
    cursor = 0
    stream_index = 0
    for lo in range(0, verts, chunk):
        n = min(chunk, verts - lo)
        v = P.gen_c2(first_vertex + lo, n, threads).view(np.uint8).reshape(-1)
        firsts = np.arange(0, n, segment, dtype=np.uint64)
        cnts = np.minimum(np.uint64(segment), np.uint64(n) - firsts).astype(np.uint64)
        blob, o, s = R.encode_segments(v, 32, firsts, cnts, level, version, threads)
        blob = blob[: int(o[-1] + ((s[-1] + 15) & ~np.uint64(15)))]
        blobs.append(blob)
        offs.append(o + np.uint64(cursor))
        sizes.append(s)
        counts.append(cnts)
        for j in (0, len(firsts) // 2, len(firsts) - 1):
            a, c = int(firsts[j]) * 32, int(cnts[j]) * 32
            check[stream_index + j] = v[a : a + c].copy()
        cursor += blob.size
        stream_index += len(firsts)
    blob = np.concatenate(blobs + [np.zeros(64, np.uint8)])
    offsets, sizes, counts = np.concatenate(offs), np.concatenate(sizes), np.concatenate(counts)
    return dict(blob=blob, offsets=offsets, sizes=sizes, counts=counts, check=check,
                decoded_bytes=int(counts.sum()) * 32, encoded_bytes=int(sizes.sum()))


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

def cpu_decode(wl, passes: int, threads: int, max_streams=None):
    from oracle import loader

    lib = loader.ref() if loader.have_ref() else loader.port()
    n = len(wl["offsets"]) if max_streams is None else min(max_streams, len(wl["offsets"]))
    streams = []
    for i in range(n):
        o, s = int(wl["offsets"][i]), int(wl["sizes"][i])
        streams.append((wl["blob"][o : o + s], int(wl["counts"][i]), 32, 0))
    best, times, outs, status = lib.decode_batch_mt(streams, threads, passes)
    assert all(s == 0 for s in status)
    decoded = sum(int(wl["counts"][i]) * 32 for i in range(n))
    return dict(seconds=best, times=times, decoded_bytes=decoded, kind=lib.kind, threads=threads, n_streams=n)


def run_reference_arm(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    from oracle import loader

    lib = loader.ref() if loader.have_ref() else loader.port()
    threads = lib.hw_threads()
    wl = build_workload(args.verts, args.segment, args.level, args.version, 0, threads)
    r_warm = cpu_decode(wl, max(1, args.warmup), threads)
    r = cpu_decode(wl, args.steps, threads)
    per_step = float(np.mean(r["times"]))
    value = wl["decoded_bytes"] / per_step / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, wl),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": r["kind"],
                         "sample": f"full workload, {r['n_streams']} streams, one meshopt_decodeVertexBuffer call per stream, {threads} host threads, mean of {args.steps} passes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, wl):
    n = len(wl["offsets"])
    return {
        "workload": f"BASELINE configs[1]: v1 codec, encode level {args.level}, {args.verts} vertices x 32 bytes per GPU "
                    f"({wl['decoded_bytes']/1e9:.3f} GB decoded, {wl['encoded_bytes']/1e9:.3f} GB encoded), reference-encoded as {n} independent "
                    f"streams of {args.segment} vertices",
        "vertex_size": 32, "vertices_per_gpu": args.verts, "streams_per_gpu": n, "segment_vertices": args.segment,
        "codec_version": args.version, "encode_level": args.level,
        "l2_policy": "no flush: encoded+decoded working set per step (>= 3 GB) is far larger than the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------
# main GPU arm
# ------------------------------------------------------------------------------------------------

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

    host_threads = max(1, (os.cpu_count() or 1) // max(1, world))
    t_gen = time.time()
    wl = build_workload(args.verts, args.segment, args.level, args.version, rank * args.verts, host_threads)
    t_gen = time.time() - t_gen
    n = len(wl["offsets"])

    # ---- device-resident arm ---------------------------------------------------------------------
    ctx = mb.Context(local)
    blob = torch.from_numpy(wl["blob"]).to(dev)
    out_lens = (wl["counts"] * np.uint64(32) + np.uint64(15)) & ~np.uint64(15)
    out_offs = np.zeros(n, np.uint64)
    np.cumsum(out_lens[:-1], out=out_offs[1:])
    out = torch.empty(int(out_lens.sum()) + 64, dtype=torch.uint8, device=dev)
    items = [(blob.data_ptr() + int(wl["offsets"][i]), int(wl["sizes"][i]), out.data_ptr() + int(out_offs[i]), int(wl["counts"][i]), 32, 0) for i in range(n)]
    plan = mb.Plan(ctx, mb.make_streams(items))
    plan_launches = plan.launches
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        plan.run(stream)
    status = plan.status(stream)
    assert (status == 0).all(), "decode reported errors"
    host_out = None
    for si, want in wl["check"].items():  # parity spot check against the original vertices (outside the timed region)
        o = int(out_offs[si])
        got = out[o : o + want.size].cpu().numpy()
        assert np.array_equal(got, want), f"stream {si} decoded incorrectly"

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        plan.run(stream)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * wl["decoded_bytes"] / (ms_per_step * 1e-3) / 1e9

    hist = plan.timing_history(min(64, args.steps))
    decode_ms = float(np.mean([h["decode_ms"] for h in hist]))
    walk_ms = float(np.mean([h["walk_ms"] for h in hist]))
    alg_bytes = wl["encoded_bytes"] + wl["decoded_bytes"]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (decode_ms * 1e-3) / 1e9
    # DRAM bytes of one launch of this workload, from the committed ncu --set full capture (profiles/): not measured live
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        for rec in json.load(open(tpath)):
            if rec.get("verts") == args.verts and rec.get("segment") == args.segment and rec.get("level") == args.level and rec.get("version") == args.version:
                traffic, traffic_src = rec["dram_bytes_per_launch"], rec.get("source")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                "kernel": "decode_kernel", "kernel_ms": decode_ms, "walk_kernel_ms": walk_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_in = torch.from_numpy(wl["blob"]).pin_memory()
        h_out = torch.empty(int(out_lens.sum()) + 64, dtype=torch.uint8).pin_memory()
        hitems = [(h_in.data_ptr() + int(wl["offsets"][i]), int(wl["sizes"][i]), h_out.data_ptr() + int(out_offs[i]), int(wl["counts"][i]), 32, 0) for i in range(n)]
        harr = mb.make_streams(hitems)
        del blob, out, plan
        torch.cuda.empty_cache()
        rc = mb.lib().mob200_decode_batch_host(ctx.handle, harr, n)  # warm-up (allocates device arenas)
        assert rc == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rc = mb.lib().mob200_decode_batch_host(ctx.handle, harr, n)
            assert rc == 0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_s = (t1 - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        for si, want in wl["check"].items():
            o = int(out_offs[si])
            assert np.array_equal(h_out[o : o + want.size].numpy(), want), f"e2e: stream {si} decoded incorrectly"
        e2e = {"value": world * wl["decoded_bytes"] / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": wl["encoded_bytes"], "d2h_bytes_per_step": wl["decoded_bytes"],
               "ms_per_step": e2e_s * 1e3, "api": "mob200_decode_batch_host (pinned host buffers)"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = (loader.ref() if loader.have_ref() else loader.port()).hw_threads()
        r = cpu_decode(wl, args.cpu_passes, threads)
        cpu = {"value": r["decoded_bytes"] / r["seconds"] / 1e9, "unit": UNIT, "cores": threads, "kind": r["kind"],
               "sample": f"full workload ({r['n_streams']} streams, {r['decoded_bytes']/1e9:.2f} GB decoded), best of {args.cpu_passes} passes, {threads} host threads"}
        r1 = cpu_decode(wl, 1, 1, max_streams=max(1, (1 << 23) // args.segment))
        cpu["single_thread_value"] = r1["decoded_bytes"] / r1["seconds"] / 1e9

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, wl), "clocks": clocks, "e2e": e2e, "gpu_launches": int(args.steps * plan_launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "notes": {"generation_seconds": t_gen, "kernels_per_step": ["decode_kernel (one persistent kernel: walker, producer and decoder warps)"]},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
