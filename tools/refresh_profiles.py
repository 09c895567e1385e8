"""Copy one GPU pass (tools/gpu_round2.sh, TAG=...) from gpurun_out/ into profiles/r2_* and rebuild the two summaries
that are derived from it (ncu summary of the headline kernel; SASS mnemonic counts of the built library).
    python tools/refresh_profiles.py r2v"""
import json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in [("bench.json", "r2_bench.json"), ("bench_reference.json", "r2_bench_reference.json"), ("meshlet.json", "r2_meshlet.json"),
                 ("filters.json", "r2_filters.json"), ("index.json", "r2_index.json"), ("launches.csv", "r2_launches_bench.csv"),
                 ("gpu.txt", "r2_gpu.txt"), ("lscpu.txt", "r2_host_lscpu.txt")]:
    s = os.path.join(G, f"{tag}_{src}")
    if os.path.exists(s) and os.path.getsize(s):
        shutil.copy(s, os.path.join(P, dst))
    else:
        print("missing", s)

rep = os.path.join(G, f"{tag}_full.ncu-rep")
if os.path.exists(rep):
    run = lambda *a: subprocess.run([sys.executable, *a], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=ROOT).stdout
    b = json.load(open(os.path.join(G, f"{tag}_bench.json")))
    head = (f"# ncu --set full --clock-control none, decode_kernel<0,0,1> (block mode: fused walker + producer + decoder; one walker lane per block from "
            f"the sidecar), headline workload: {b['config']['workload'][:160]}..., one launch (tools/gpu_round2.sh, TAG={tag}, kernel sources {bench.source_hash()})\n")
    txt = head + run("tools/ncu_summary.py", rep)
    txt += "\n# hottest source lines (share of executed warp-instructions in ncu's instrumented pass / of warp stall samples)\n" + run("tools/ncu_lines.py", rep, "30")
    txt += "\n# instruction footprint\n" + run("tools/ncu_hotcode.py", rep)
    open(os.path.join(P, "r2_ncu_decode_kernel_summary.txt"), "w").write(txt)

lib = os.path.join(ROOT, "meshoptimizer_b200", "lib", "libmeshopt_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
kern, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        kern[cur].append(line)
cols = ["UBLKCP", "SYNCS", "LDGSTS", "PRMT", "IDP", "SHFL", "LDS", "STS"]
out = ["# SASS mnemonic counts per kernel of meshoptimizer_b200/lib/libmeshopt_b200.so (cuobjdump -sass; sm_100a only), kernel sources " + bench.source_hash()]
for k in sorted(kern):
    body = kern[k]
    cnt = [sum(1 for l in body if re.search(r"\b" + c + r"\b|\b" + c + r"\.", l)) for c in cols]
    tensor = sum(1 for l in body if re.search(r"UTCMMA|UTMALDG|HMMA|UTCHMMA", l))
    out.append(f"{k:70s} {len(body):6d} " + " ".join(f"{c:6d}" for c in cnt) + f" {tensor}")
out.append(f"{'kernel':70s} {'instr':>6s} " + " ".join(f"{c:>6s}" for c in cols) + " tensor/TMA-tile")
out.append("\n# arch of every cubin in the library:")
out += ["  " + l.strip() for l in subprocess.run(["cuobjdump", "-lelf", lib], stdout=subprocess.PIPE, text=True).stdout.splitlines()]
hk = next((k for k in kern if "decode_kernelILb0ELb0ELb1E" in k), None)
if hk:
    out.append("\n# excerpt: mbarrier (SYNCS) and TMA bulk-copy (UBLKCP) instructions of decode_kernel<0,0,1>, the headline instantiation")
    out += [l.rstrip() for l in kern[hk] if "UBLKCP" in l or "SYNCS" in l][:24]
    out.append("\n# excerpt: byte-permute / dot-product inner loop of the decoder warps (first PRMT / IDP instructions)")
    out += [l.rstrip() for l in kern[hk] if "PRMT" in l or "IDP" in l][:12]
open(os.path.join(P, "r2_sass_excerpt.txt"), "w").write("\n".join(out) + "\n")
print("profiles refreshed from", tag, "sources", bench.source_hash())
