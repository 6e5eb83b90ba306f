"""bench.py's own arm (N = 1) on the CPU-emulated device of tests/emu with a stand-in `torch`: a dry run of the
script's control flow before it is spent on the B200.  The numbers it prints describe the emulation, not the GPU.
Developer tool.

    python tools/bench_dry_run.py --steps 2 --warmup 3 --cpu-sample-pods 16
"""
import os, sys, types, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import build_emu_cuda
os.environ['NHD_B200_LIB']=build_emu_cuda.build()
os.environ['NHD_B200_ALLOW_EMULATED'] = '1'; os.environ['EMU_LANE_ORDER']='d'
import numpy as np
class FakeTensor:
    def __init__(self, n): self.a = np.zeros(1, dtype=np.uint8)
    def zero_(self): return self
torch = types.ModuleType('torch')
torch.uint8 = 'uint8'
torch.empty = lambda n, dtype=None, device=None: FakeTensor(n)
torch.cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda d: None, synchronize=lambda: None)
torch.device = lambda *a: None
dist = types.ModuleType('torch.distributed')
torch.distributed = dist
sys.modules['torch'] = torch; sys.modules['torch.distributed'] = dist
sys.argv = ['bench.py'] + sys.argv[1:]
os.chdir(ROOT)
runpy.run_path(os.path.join(ROOT, 'bench.py'), run_name='__main__')
