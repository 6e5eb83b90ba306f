"""Workload for compute-sanitizer (racecheck / synccheck / memcheck): small clusters through every sweep mode —
standing decisions with the two pod classes side by side (two CTAs), on one warp, without standing decisions,
moving clocks, spills — each compared with the oracle.  Sized to finish under the sanitizer's ~100x slowdown."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workload
from nhd_b200.solver import Solver
from oracle import binding
from tests import helpers, ref_compare, scenarios

bad = 0
runs = 0


def check(recs, speed, pods, now, min_busy=30.0, **mode):
    global bad, runs
    s = Solver(speed, min_busy_secs=min_busy, **mode)
    try:
        s.load_nodes(recs)
        got = s.solve_batch(pods, now)
        final = s.read_nodes()
    finally:
        s.close()
    want, wf = binding.solve(recs, speed, pods, now, min_busy_secs=min_busy)
    ok = helpers.binding_bytes_equal(want, got) and final.tobytes() == wf.tobytes()
    runs += 1
    bad += not ok


for cfg, nn, npods in ((3, 2048, 160), (5, 1024, 96), (2, 512, 64)):
    recs, speed, pods, now = workload.make_workload(cfg, n_nodes=nn, n_pods=npods)
    for mode in (dict(), dict(cpu_warps=1), dict(sweep_debug=1), dict(single_warp=True)):
        check(recs, speed, pods, now, **mode)
    rng = np.random.default_rng(cfg)
    moving = 1000.0 + np.cumsum(rng.choice([0.0, 0.0, 5.0, 31.0, -40.0], size=len(pods)))
    check(recs, speed, pods, moving)
for seed in range(6):
    scn = scenarios.random_scenario(8000 + seed * 3, n_nodes=12, n_pods=48, flavor='mixed' if seed % 2 else 'big', max_groups=3)
    recs, pods, _, layout = ref_compare.pack_scenario(scn)
    now = np.full(len(pods), 5000.0)
    for busy in (30.0, 0.0):
        for mode in (dict(), dict(cpu_warps=1)):
            check(recs, layout.speed_table(), pods, now, min_busy=busy, **mode)
print(f'sanitizer workload: {runs} batches, {bad} differ from the oracle')
sys.exit(1 if bad else 0)
