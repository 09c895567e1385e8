"""Debug helper: run the C4 small-stream workload and list the streams whose status / bytes are wrong."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meshoptimizer_b200 as mb
from oracle import loader, workloads
from tests.gpu_util import device_run

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
w = workloads.c4(n)
want = workloads.expected_outputs(w, lib=loader.ref())
for attempt in range(3):
    outs, status, plan, guard = device_run(w)
    bad = [i for i in range(w.n) if status[i] != 0 or not np.array_equal(outs[i], want[i])]
    print(f"attempt {attempt}: guard={guard} bad streams {len(bad)} of {w.n}; status histogram", dict(zip(*np.unique(status, return_counts=True))))
    for i in bad[:12]:
        d = np.nonzero(outs[i] != want[i])[0]
        print("  stream", i, "status", int(status[i]), "count", int(w.counts[i]), "vs", int(w.vertex_sizes[i]), "size", int(w.sizes[i]), "src&15", int(w.offsets[i]) & 15, "first bad byte", int(d[0]) if d.size else None, "bad bytes", int(d.size))
