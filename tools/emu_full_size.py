"""A whole BASELINE configuration through the real kernels on the CPU-emulated device (tests/emu), every binding
and every final node record compared with the (multi-threaded) oracle.  Developer tool; needs no GPU.

    python tools/emu_full_size.py 4        # 65 536 nodes x 4 096 pods: ~4 s emulated device + ~75 s oracle on 8 cores
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import build_emu_cuda
os.environ['NHD_B200_LIB']=build_emu_cuda.build()
os.environ['NHD_B200_ALLOW_EMULATED'] = '1'; os.environ['EMU_LANE_ORDER']='d'
import workload
from nhd_b200.solver import Solver
from oracle import binding
from tests import helpers
cfg=int(sys.argv[1])
recs, speed, pods, now = workload.make_workload(cfg)
print(len(recs), len(pods), flush=True)
s=Solver(speed); s.load_nodes(recs)
t0=time.time(); b=s.solve_batch(pods, now); print('emulated device', time.time()-t0,'s', flush=True)
final=s.read_nodes(); s.close()
t0=time.time(); ob, orecs = binding.solve(recs, speed, pods, now, threads=os.cpu_count()); print('oracle mt', time.time()-t0,'s', flush=True)
print('bindings equal', helpers.binding_bytes_equal(ob,b), 'final equal', final.tobytes()==orecs.tobytes(), 'placed', int((ob['status']==0).sum()))
