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
names = ['pod header', 'bitmap scan', 'summary+memo lookup', 'full eval (miss)', 'decode/exit', 'assign_resources', 'write-back']
tot = sum(int(x) for x in c[:16])
print(f'cfg{cfg}: sweep {t["sweep_ms"]:.3f} ms, {len(pods)} pods, total cycles {tot} ({tot/len(pods):.0f}/pod)')
for i, n in enumerate(names):
    print(f'  {n:24s} cycles {int(c[i]):10d} ({100*int(c[i])/max(tot,1):5.1f}%)  count {int(c[32+i]):7d}  avg {int(c[i])/max(int(c[32+i]),1):8.1f}')
print('  stale candidates', int(c[32 + 8]))
# CPU-class warp 0 only: pods it worked out ahead / results that went stale while waiting / results adopted
print('  warp 0: speculations', int(c[32 + 15]), 'invalidated', int(c[32 + 14]), 'adopted', int(c[32 + 13]))
for i, n in enumerate(['cpu2: B mask + layout', 'cpu2: sub-problems', 'cpu2: combine', 'cpu2: mapping memo', 'cpu2: expand', 'cpu2: claim order']):
    print(f'  {n:24s} cycles {int(c[16+i]):10d}  per CPU pod {int(c[16+i])/2048:8.1f}')
s.close()
