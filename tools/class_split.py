"""Developer probe: sweep time of the CPU-only pods and of the GPU pods of a config, each on their own."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workload
from nhd_b200.solver import Solver

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
recs, speed, pods, now = workload.make_workload(cfg)
is_gpu = pods['groups']['n_gpus'].sum(axis=1) > 0
for name, sel in (('cpu-only pods', ~is_gpu), ('gpu pods', is_gpu), ('all', np.ones(len(pods), bool))):
    for cw in (1, 3, 5):
        s = Solver(speed, cpu_warps=cw)
        s.load_nodes(recs); s.snapshot()
        best = 1e9
        for it in range(4):
            s.restore()
            s.solve_batch(pods[sel], now[sel])
            best = min(best, s.timing()['sweep_ms'])
        print(f'cfg{cfg} {name:14s} n={int(sel.sum())} cpu_warps={cw} sweep={best:.3f} ms', flush=True)
        s.close()
