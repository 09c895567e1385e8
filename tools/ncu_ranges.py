"""Share of executed warp-instructions / stall samples per source-line range of one file of an ncu report:
   python tools/ncu_ranges.py report.ncu-rep file.cuh name:lo:hi ..."""
import csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = [tuple(x.split(':')) for x in sys.argv[3:]]
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
        agg[(cur, int(r[0]))] = (num(r[4]), num(r[7]))
ti = sum(v[1] for v in agg.values()) or 1; ts = sum(v[0] for v in agg.values()) or 1
print('total inst', ti, 'samples', ts)
byfile = {}
for (f, l), (s, i) in agg.items():
    byfile.setdefault(f, [0, 0]); byfile[f][0] += s; byfile[f][1] += i
for f, (s, i) in byfile.items():
    print(f, 'inst %.1f%% samp %.1f%%' % (100 * i / ti, 100 * s / ts))
for name, lo, hi in ranges:
    lo = int(lo); hi = int(hi)
    s = sum(v[0] for (f, l), v in agg.items() if f == fname and lo <= l <= hi)
    i = sum(v[1] for (f, l), v in agg.items() if f == fname and lo <= l <= hi)
    print('%-16s lines %d-%d inst %.1f%% samp %.1f%%' % (name, lo, hi, 100 * i / ti, 100 * s / ts))
