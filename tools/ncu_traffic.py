"""Read dram__bytes_{read,write}.sum of the decode kernel from an .ncu-rep and record it in profiles/dram_traffic.json
together with the hash of the kernel sources it was captured on (bench.py quotes it only for that build).

    python tools/ncu_traffic.py report.ncu-rep <config name> [verts] [level] [version]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rep, config = sys.argv[1], sys.argv[2]
verts = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 26
level = int(sys.argv[4]) if len(sys.argv) > 4 else 2
version = int(sys.argv[5]) if len(sys.argv) > 5 else 1
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
vals = {}
for r in rows[2:]:
    if len(r) != len(hdr) or "decode_kernel" not in r[hdr.index("Kernel Name")]:
        continue
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
                 "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        if name in hdr:
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(u, 1.0)
            vals.setdefault(name, []).append(v * scale)
assert vals, "no decode_kernel launch in the report"
mean = {k: sum(v) / len(v) for k, v in vals.items()}
rec = {"config": config, "verts": verts, "level": level, "version": version, "source_hash": bench.source_hash(),
       "dram_bytes_read": mean["dram__bytes_read.sum"], "dram_bytes_write": mean["dram__bytes_write.sum"],
       "dram_bytes_per_launch": mean["dram__bytes_read.sum"] + mean["dram__bytes_write.sum"], "kernel_ms_under_ncu": mean.get("gpu__time_duration.sum"),
       "warp_instructions": mean.get("smsp__inst_executed.sum"), "issue_pct": mean.get("sm__inst_issued.avg.pct_of_peak_sustained_active"),
       "warps_active_pct": mean.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "dram_pct": mean.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
       "source": f"ncu --set full, {os.path.basename(rep)}, mean of {len(vals['dram__bytes_read.sum'])} launch(es)"}
path = os.path.join(ROOT, "profiles", "dram_traffic.json")
recs = json.load(open(path)) if os.path.exists(path) else []
recs = [r for r in recs if not (r.get("config") == config and r.get("verts") == verts and r.get("level") == level and r.get("version") == version)]
recs.append(rec)
json.dump(recs, open(path, "w"), indent=1)
print(json.dumps(rec, indent=1))
