"""Static SASS size of the lane-walker decode kernel per source region (nvdisasm -g line info).
   python tools/sass_regions.py   (run after a build)"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "meshoptimizer_b200", "lib", "libmeshopt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith("mob200_kernels.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
src = open(os.path.join(ROOT, "meshoptimizer_b200", "csrc", "mob200_decoder.cuh")).read().splitlines()
def find(pat):
    for i, l in enumerate(src):
        if pat in l: return i + 1
    return 10**9
marks = [("helpers", 1), ("producer", find("__device__ void producer_main")), ("unpack", find("uint4 unpack_group(")), ("decoder_head", find("__device__ void decoder_main")),
         ("dec_items", find("for (uint32_t base = warp_base;")), ("dec_store", find("this warp no longer needs the slot"))]
cur_fn = None; cur = None; counts = {}
for line in dis.splitlines():
    m = re.match(r'\.text\.(\S+):', line)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', line) and cur_fn and 'ILb0' in cur_fn and cur:
        f, l = cur
        if f == 'mob200_decoder.cuh':
            k = [name for name, start in marks if l >= start][-1]
        else:
            k = f
        counts[k] = counts.get(k, 0) + 1
for k, v in sorted(counts.items(), key=lambda kv: -kv[1]):
    print(f"{k:28s} {v:5d} instr {v*16/1024:5.1f} KB")
print('total', sum(counts.values()))
