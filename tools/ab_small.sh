for so in meshoptimizer_b200/lib/variants/*.so; do echo "== $(basename $so .so)"; MOB200_LIB=$PWD/$so timeout 200 python tools/bench_small_vs.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  vs',d['vertex_size'],d['streams'],'%.4f ms %.0f GB/s'%(d['best_ms'],d['decoded_GBps']),d['ok'])"; done
