"""Device-resident timing of the bench workload without parity checks (diagnostics: used with the
MOB200_* environment knobs, e.g. MOB200_WALKER_LEAD=4294967295 = walkers only).

    python tools/gpu_diag.py [verts] [segment] [runs]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import meshoptimizer_b200 as mb
import bench

verts = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
segment = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 12
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 10
check = os.environ.get("MOB200_WALKER_LEAD") is None

wl = bench.build_workload(verts, segment, 2, 1, 0, os.cpu_count() or 1)
n = len(wl["offsets"])
dev = torch.device("cuda:0")
ctx = mb.Context(0)
blob = torch.from_numpy(wl["blob"]).to(dev)
out_lens = (wl["counts"] * np.uint64(32) + np.uint64(15)) & ~np.uint64(15)
out_offs = np.zeros(n, np.uint64)
np.cumsum(out_lens[:-1], out=out_offs[1:])
out = torch.empty(int(out_lens.sum()) + 64, dtype=torch.uint8, device=dev)
items = [(blob.data_ptr() + int(wl["offsets"][i]), int(wl["sizes"][i]), out.data_ptr() + int(out_offs[i]), int(wl["counts"][i]), 32, 0) for i in range(n)]
plan = mb.Plan(ctx, mb.make_streams(items))
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    plan.run(stream)
plan.debug_counters(reset=True)
for _ in range(runs):
    plan.run(stream)
torch.cuda.synchronize()
dbg = plan.debug_counters(reset=True)
hist = plan.timing_history(runs)
ms = sorted(h for h in hist)
ok = None
if check:
    status = plan.status(stream)
    ok = bool((status == 0).all())
    for si, want in wl["check"].items():
        o = int(out_offs[si])
        ok = ok and np.array_equal(out[o : o + want.size].cpu().numpy(), want)
alg = wl["encoded_bytes"] + wl["decoded_bytes"]
dt, pt = max(1, dbg["decoder_total"]), max(1, dbg["producer_total"])
print("  decoders: wait full %.1f%% carry %.1f%% tile %.1f%% | producers: meta %.1f%% wait-slot %.1f%% lookback %.1f%% | decoder cycles per CTA-run %.0f" % (
    100 * dbg["decoder_wait_full"] / dt, 100 * dbg["decoder_wait_carry"] / dt, 100 * dbg["decoder_wait_tile"] / dt,
    100 * dbg["producer_meta"] / pt, 100 * dbg["producer_wait_slot"] / pt, 100 * dbg["producer_lookback"] / pt, dt / runs / max(1, plan.grid if hasattr(plan, "grid") else 740)))
wt = max(1, dbg["walker_total"])
print("  walkers: wait-prev %.1f%% wait-hard %.1f%% | refills %d hard waits %d | cycles per refill %.0f" % (100 * dbg["walker_wait_prev"] / wt, 100 * dbg["walker_wait_hard"] / wt, dbg["walker_refills"] // runs, dbg["walker_hard_waits"] // runs, wt / max(1, dbg["walker_refills"])))
print("  walkers: steps %.1f%% refills %.1f%% of walker time | general channels %d per run | walker cycles per warp-run %.0f" % (100 * dbg["r13"] / wt, 100 * dbg["r14"] / wt, dbg["r15"] // runs, wt / runs / 512))
print(f"verts {verts} segment {segment} streams {n} env {{{', '.join(k + '=' + v for k, v in os.environ.items() if k.startswith('MOB200_'))}}}: "
      f"best {ms[0]:.3f} ms median {ms[len(ms)//2]:.3f} ms | decoded {wl['decoded_bytes']/ms[0]/1e6:.0f} GB/s traffic {alg/ms[0]/1e6:.0f} GB/s ({alg/ms[0]/1e6/6543.4:.3f} of peak) parity={ok}", flush=True)
