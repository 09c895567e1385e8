"""Instruction-cache view of an ncu report: how many distinct SASS instructions carry the executed instructions.
   python tools/ncu_hotcode.py report.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; ins = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr):
        try:
            addr = int(r[0], 16) if r[0].startswith('0x') else int(r[0])
        except Exception:
            continue
        ex = r[hdr.index('Instructions Executed')]
        ins.append((addr, int(ex) if ex.isdigit() else 0, r[1]))
if not ins:
    print('no sass rows; header', hdr); sys.exit()
tot = sum(e for _, e, _ in ins)
print('SASS instructions', len(ins), '=', len(ins) * 16 / 1024, 'KB; executed warp-instructions', tot)
for frac in (0.5, 0.8, 0.9, 0.95, 0.99):
    acc = 0; n = 0
    for a, e, _ in sorted(ins, key=lambda x: -x[1]):
        acc += e; n += 1
        if acc >= frac * tot: break
    print(f'  {frac*100:.0f}% of executed instructions come from {n} SASS instructions = {n*16/1024:.1f} KB')
# 4 KB windows
win = {}
for a, e, _ in ins:
    win.setdefault(a // 4096, [0, 0]); win[a // 4096][0] += e; win[a // 4096][1] += 1
print('executed share per 4 KB of code:')
print('  ' + ' '.join(f'{100*v[0]/tot:.0f}' for k, v in sorted(win.items())))
