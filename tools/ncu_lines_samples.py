"""Like ncu_lines.py, sorted by warp stall samples:  python tools/ncu_lines_samples.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; agg = {}; hdr = None
def num(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and r[0].strip().isdigit():
        agg[(cur, int(r[0]))] = (num(r[4]), num(r[7]), r[1].strip())
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
print('total samples', ts, 'total warp-instructions', ti)
for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} inst {i/ti*100:5.1f}% samp {s/ts*100:5.1f}%  {src[:110]}")
