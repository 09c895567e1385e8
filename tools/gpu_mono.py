import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import workloads
from tests.gpu_util import device_run
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = workloads.c2(total=total, seg=seg or None)
outs, status, plan, guard = device_run(w, runs=2)
print(w.name, plan.last_timing(), np.array_equal(np.concatenate(outs), w.source))
