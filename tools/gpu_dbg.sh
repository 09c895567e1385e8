MOB200_DEBUG_COUNTERS=1 python -c "from meshoptimizer_b200 import build; build.build(force=True)"
python tools/gpu_diag.py 67108864 4096 10
MOB200_WALKER_LEAD=4294967293 python tools/gpu_diag.py 67108864 4096 10
python -c "from meshoptimizer_b200 import build; build.build(force=True)"
MOB200_WALKER_LEAD=4294967293 python tools/gpu_diag.py 67108864 4096 10 | tail -1
