"""Developer timing probe (not the contract bench): per-config kernel timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workload
from nhd_b200.solver import Solver

CW = int(os.environ.get('NHD_CPU_WARPS', '0'))
for cfg in [int(a) for a in sys.argv[1:]] or [2, 3, 4]:
    recs, speed, pods, now = workload.make_workload(cfg)
    s = Solver(speed, cpu_warps=CW)
    s.load_nodes(recs)
    s.snapshot()
    best = None
    for it in range(4):
        s.restore()
        t0 = time.perf_counter()
        b = s.solve_batch(pods, now)
        wall = (time.perf_counter() - t0) * 1e3
        t = s.timing()
        if best is None or t['total_ms'] < best['total_ms']:
            best = dict(t, wall_ms=wall)
    print(f"cw={CW} cfg{cfg} N={len(recs)} P={len(pods)} placed={int((b['status']==0).sum())} types={best['n_types']} "
          f"filter={best['filter_ms']:.3f}ms sweep={best['sweep_ms']:.3f}ms total={best['total_ms']:.3f}ms wall={best['wall_ms']:.3f}ms "
          f"-> {len(pods)/best['total_ms']*1e3:,.0f} dec/s (device) {len(pods)/best['wall_ms']*1e3:,.0f} dec/s (e2e wall)", flush=True)
    s.close()
