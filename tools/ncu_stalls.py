"""Per source line: where the samples of one stall reason fall.  python tools/ncu_stalls.py report.ncu-rep stall_no_inst [top]"""
import csv, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        try: v = int(r[hdr.index(col)])
        except Exception: v = 0
        agg[(cur, int(r[0]))] = (v, r[1].strip())
tot = sum(v for v, _ in agg.values()) or 1
print(col, 'total samples', tot, '| columns with stall_:', [h for h in hdr if h.startswith('stall_')])
for (f, l), (v, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:4d} {100*v/tot:5.1f}%  {src[:100]}")
