"""One staged solve of a config (for ncu): python tools/profile_run.py <config> [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workload
from nhd_b200.solver import Solver
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
recs, speed, pods, now = workload.make_workload(cfg)
s = Solver(speed, cpu_warps=int(os.environ.get('NHD_CPU_WARPS', '0')))
s.load_nodes(recs)
s.snapshot()
s.stage_batch(pods, now)
for _ in range(reps):
    s.restore()
    s.solve_staged()
    s.sync()
print(s.timing())
s.close()
