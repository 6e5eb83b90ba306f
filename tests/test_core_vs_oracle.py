"""CPU-only: the product's scalar core (nhd_core.cuh compiled with g++, tests/emu) must agree
with the oracle bit for bit — bindings and final node records — on randomized clusters."""
import ctypes

import numpy as np
import pytest

from nhd_b200 import wire
from tests import helpers, ref_compare, scenarios


@pytest.mark.parametrize('two_stage', [False, True], ids=['direct', 'two_stage'])
@pytest.mark.parametrize('flavor', ['mixed', 'wild', 'vf', 'big'])
def test_core_matches_oracle_random(oracle_lib, emu, flavor, two_stage):
    placed = 0
    for seed in range(40):
        scn = scenarios.random_scenario(1000 + seed * 13 + len(flavor), n_nodes=8, n_pods=40, flavor=flavor,
                                        max_groups=4 if flavor != 'wild' else 3)
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        ob, orecs = oracle_lib.solve(recs, layout.speed_table(), pods, now)
        eb, erecs = helpers.emu_solve(emu, recs, layout.speed_table(), pods, now, two_stage=two_stage)
        assert helpers.binding_bytes_equal(ob, eb), helpers.first_binding_diff(ob, eb)
        assert orecs.tobytes() == erecs.tobytes(), ref_compare.diff_records(orecs, erecs)[:3]
        placed += int((ob['status'] == 0).sum())
    assert placed > 100


def test_snapshot_predicate_matches_oracle_candidates(oracle_lib, emu):
    """node_feasible (what the CUDA filter kernel evaluates per thread) == membership in
    filts[1] after IntersectResources, on a partially filled cluster, busy window aside."""
    checked = 0
    for seed in range(25):
        scn = scenarios.random_scenario(77 + seed, n_nodes=12, n_pods=30, flavor='wild' if seed % 2 else 'mixed')
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        _, filled = oracle_lib.solve(recs, layout.speed_table(), pods[:15], now[:15])
        for pod in pods[15:]:
            cand = oracle_lib.candidates(filled, layout.speed_table(), pod, now=1e9)   # nobody busy
            feas = helpers.emu_feasible(emu, filled, layout.speed_table(), pod)
            if pod['map_type'] in (1, 2):
                assert np.array_equal(cand, feas), (seed, cand, feas)
                checked += int(cand.sum())
    assert checked > 50


def test_register_only_predicate_equals_the_general_one(oracle_lib, emu):
    """node_feasible_k2 (2-NUMA nodes, pods of one or two groups: what the filter kernel runs for the types outside
    its tables) == node_feasible, pair by pair, on partially filled clusters of every flavour — NUMA and PCI mode,
    shared NICs, switches shared across NUMA nodes, SR-IOV VFs — and on the benchmark clusters."""
    import workload
    applied = yes = masks = entries = 0
    cases = []
    for seed in range(60):
        flavor = ('mixed', 'wild', 'vf', 'big')[seed % 4]
        scn = scenarios.random_scenario(4100 + seed * 7, n_nodes=16, n_pods=40, flavor=flavor, max_groups=3)
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        cases.append((recs, layout.speed_table(), pods, now, 20))
    for config in (3, 5):
        for wild in (False, True):
            recs, speed, pods, now = workload.make_workload(config, n_nodes=256, n_pods=600, wild=wild)
            cases.append((recs, speed, pods, now, 400))
    for recs, speed, pods, now, n_fill in cases:
        _, filled = oracle_lib.solve(recs, speed, pods[:n_fill], now[:n_fill])
        seen = set()
        for pod in pods[n_fill:]:
            key = pod.tobytes()
            if key in seen or pod['map_type'] not in (1, 2):
                continue
            seen.add(key)
            gen = helpers.emu_feasible(emu, filled, speed, pod)
            k2 = helpers.emu_feasible(emu, filled, speed, pod, k2=True)
            on = k2 != 2
            assert np.array_equal(gen[on], k2[on]), (np.flatnonzero(on & (gen != k2))[:5], pod)
            applied += int(on.sum())
            yes += int(k2[on].sum())
            # and the forms resolve_kernel uses: the three stage masks, the first NIC entry of every NUMA tuple
            cnt = (ctypes.c_int * 2)()
            pr = np.ascontiguousarray(pod, dtype=wire.POD_DTYPE).reshape(1)
            sp = np.ascontiguousarray(speed, dtype='<f8')
            fr = np.ascontiguousarray(filled, dtype=wire.NODE_DTYPE)
            emu.nhd_emu_check_k2.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]
            assert emu.nhd_emu_check_k2(0.9, sp.ctypes.data, len(fr), fr.ctypes.data, pr.ctypes.data, cnt) == 0, pod
            masks += cnt[0]
            entries += cnt[1]
    assert applied > 5000 and yes > 500 and applied - yes > 500, (applied, yes)
    assert masks > 5000 and entries > 10000, (masks, entries)


@pytest.mark.parametrize('config', [2, 3, 5])
def test_two_stage_core_on_baseline_shapes(oracle_lib, emu, config):
    """Summary-state decisions + deferred core ids (the CUDA sweep's structure) on the benchmark's
    cluster shapes, where nodes really fill up (prefix offsets, misc-core hyperthread overflow)."""
    import workload
    recs, speed, pods, now = workload.make_workload(config, n_nodes=192, n_pods=1500)
    ob, orecs = oracle_lib.solve(recs, speed, pods, now, min_busy_secs=0.0)
    eb, erecs = helpers.emu_solve(emu, recs, speed, pods, now, min_busy=0.0, two_stage=True)
    assert helpers.binding_bytes_equal(ob, eb), helpers.first_binding_diff(ob, eb)
    assert orecs.tobytes() == erecs.tobytes()
    assert (ob['status'] == 1).sum() > 50          # the cluster ran full
