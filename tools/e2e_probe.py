"""Developer probe: the end-to-end leg of bench.py (pinned host buffers -> nhd_load_nodes + nhd_solve_batch -> host
bindings) split into its two calls, on the library named by NHD_B200_LIB; `--nvml` initialises NVML and polls it from a
thread first (what bench.py's clock sampler does during the device-timed leg)."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import workload
from nhd_b200 import wire
from nhd_b200.solver import Solver, pinned_array

recs, speed, pods, now = workload.make_workload(4)
N, P = len(recs), len(pods)
s = Solver(speed)
nv = None
if '--nvml' in sys.argv:
    import bench
    smp = bench.ClockSampler(0, str(torch.cuda.get_device_properties(0).uuid))
    smp.start()
    time.sleep(0.1)
    nv = smp.stop()
pin_recs = pinned_array(N, wire.NODE_DTYPE); pin_recs[:] = recs
pin_pods = pinned_array(P, wire.POD_DTYPE); pin_pods[:] = pods
pin_now = pinned_array(P, '<f8'); pin_now[:] = now
out = pinned_array(P, wire.BINDING_DTYPE)
tl, ts = [], []
for it in range(13):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.load_nodes(pin_recs)
    t1 = time.perf_counter()
    s.solve_batch(pin_pods, pin_now, out=out)
    t2 = time.perf_counter()
    if it >= 3:
        tl.append(t1 - t0); ts.append(t2 - t1)
t = s.timing()
print(json.dumps({'lib': os.environ.get('NHD_B200_LIB', 'libnhd_b200.so'), 'nvml': nv,
                  'load_nodes_ms': round(1e3 * float(np.mean(tl)), 4), 'solve_batch_ms': round(1e3 * float(np.mean(ts)), 4),
                  'min_load_ms': round(1e3 * float(np.min(tl)), 4), 'min_solve_ms': round(1e3 * float(np.min(ts)), 4),
                  'device_total_ms': round(t['total_ms'], 4),
                  'e2e_decisions_per_s': round(P / (float(np.mean(tl)) + float(np.mean(ts))))}), flush=True)
s.close()
