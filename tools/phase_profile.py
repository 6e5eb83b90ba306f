"""Per-phase cycle breakdown of the sweep (needs the NHD_PROFILE build):
   python nhd_b200/build.py --profile && NHD_B200_LIB=libnhd_b200_prof.so python tools/phase_profile.py 4"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workload
from nhd_b200.solver import Solver
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
recs, speed, pods, now = workload.make_workload(cfg)
s = Solver(speed, cpu_warps=int(os.environ.get('NHD_CPU_WARPS', '0')))
s.load_nodes(recs); s.snapshot(); s.stage_batch(pods, now)
for _ in range(3):
    s.restore(); s.solve_staged(); s.sync()
    c = s.debug_counters()
t = s.timing()
names = ['pod header + slot commit', 'bitmap scan (ordinary)', 'summary+memo (ordinary)', 'full eval (ordinary)', 'decode/exit (ordinary)',
         'assign / stale mask', 'refresh of standing decisions']
tot = sum(int(x) for x in c[:16])
print(f'cfg{cfg}: sweep {t["sweep_ms"]:.3f} ms, {len(pods)} pods, warp-0 cycles {tot}')
for i, n in enumerate(names):
    print(f'  {n:34s} cycles {int(c[i]):10d} ({100*int(c[i])/max(tot,1):5.1f}%)  count {int(c[32+i]):7d}  avg {int(c[i])/max(int(c[32+i]),1):8.1f}')
print('  stale candidates (ordinary path)', int(c[32 + 8]), ' direct commits', int(c[32 + 13]), ' deferred GPU commits', int(c[32 + 12]))
print('  refresh: evaluation passes', int(c[48]), ' lanes evaluating', int(c[49]), ' passes with lanes moving on', int(c[50]),
      ' lanes moving on', int(c[51]))
s.close()
