"""Differential fuzz of the CUDA solver against the oracle on the CPU-emulated device (tests/emu):
random clusters and pod streams, every sweep mode.  Developer tool; needs no GPU.

    python tools/emu_fuzz.py [first_seed] [n_seeds]

Environment: EMU_SCHED_SEED=<n> shuffles the warp schedule, EMU_FUZZ_MODES="0,3,5" picks sweep modes, EMU_LIB_OVERRIDE=<.so>
runs another emulated build (e.g. the sources compiled with -DNHD_CHECKS, which re-derives every result the sweep adopted
ahead of its turn and traps on a difference: clean under shuffled schedules).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import build_emu_cuda                                   # noqa: E402
os.environ['NHD_B200_LIB'] = os.environ.get('EMU_LIB_OVERRIDE') or build_emu_cuda.build()
os.environ['NHD_B200_ALLOW_EMULATED'] = '1'
os.environ.setdefault('EMU_LANE_ORDER', 'd')

import numpy as np                                      # noqa: E402
import workload                                         # noqa: E402
from nhd_b200.solver import Solver                      # noqa: E402
from oracle import binding                              # noqa: E402
from tests import helpers, ref_compare, scenarios       # noqa: E402

MODES = [dict(), dict(single_warp=True), dict(cpu_warps=1), dict(sweep_debug=1), dict(sweep_debug=5), dict(cpu_warps=1), dict()]
if os.environ.get('EMU_FUZZ_MODES'):                    # e.g. "1,2": only the one-warp sweep and one CPU-class warp
    MODES = [MODES[int(i)] for i in os.environ['EMU_FUZZ_MODES'].split(',')]


def one(recs, speed, pods, now, min_busy, mode, tag):
    s = Solver(speed, min_busy_secs=min_busy, **mode)
    try:
        s.load_nodes(recs)
        half = len(pods) // 2
        b1 = s.solve_batch(pods[:half], now[:half])          # two batches: the second starts from committed state
        b2 = s.solve_batch(pods[half:], now[half:])
        final = s.read_nodes()
    finally:
        s.close()
    ob, orecs = binding.solve(recs, speed, pods, now, min_busy_secs=min_busy)
    got = np.concatenate([b1, b2])
    if not helpers.binding_bytes_equal(ob, got) or final.tobytes() != orecs.tobytes():
        print('MISMATCH', tag, mode, helpers.first_binding_diff(ob, got), flush=True)
        return False
    return True


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    bad = total = 0
    t0 = time.time()
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        flavor = ('mixed', 'wild', 'vf', 'big')[seed % 4]
        scn = scenarios.random_scenario(90000 + seed, n_nodes=int(rng.integers(3, 70)), n_pods=int(rng.integers(10, 160)),
                                        flavor=flavor, max_groups=4 if flavor != 'wild' else 3,
                                        min_busy_secs=float(rng.choice([30.0, 0.0])))
        if seed % 6 == 1:                                     # objects near the packed layout's limits
            scn = scenarios.huge_scenario(70000 + seed, float(rng.choice([30.0, 0.0])))
            flavor = 'huge'
        elif seed % 3:
            scn['now'] = [1000.0] * len(scn['now'])           # constant clock: the multi-warp sweep
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        mode = MODES[seed % len(MODES)]
        total += 1
        bad += not one(recs, layout.speed_table(), pods, now, scn['min_busy_secs'], mode, ('scenario', seed, flavor))
        if seed % 5 == 0:                                     # the benchmark's cluster shapes, nodes really fill up
            cfg = (2, 3, 4, 5)[(seed // 5) % 4]
            r, sp, p, nw = workload.make_workload(cfg, n_nodes=int(rng.integers(64, 400)), n_pods=int(rng.integers(200, 900)))
            total += 1
            bad += not one(r, sp, p, nw, float(rng.choice([30.0, 0.0])), MODES[(seed // 5) % len(MODES)], ('workload', seed, cfg))
    print(f'{total} runs, {bad} mismatches, {time.time() - t0:.0f} s')


if __name__ == '__main__':
    main()
