"""The multi-threaded CPU baseline (``nhd_oracle_solve_mt``: each pod's walk over the nodes split over
host threads, pods still strictly in order) must give the plain oracle's result bit for bit."""
import pytest

from tests import helpers, ref_compare, scenarios


@pytest.mark.parametrize('threads', [2, 3, 8, 33])
def test_mt_oracle_equals_plain_oracle_random(oracle_lib, threads):
    for seed in range(12):
        flavor = ('mixed', 'wild', 'vf', 'big')[seed % 4]
        scn = scenarios.random_scenario(4000 + seed, n_nodes=5 + 7 * (seed % 5), n_pods=40, flavor=flavor)
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        b1, r1 = oracle_lib.solve(recs, layout.speed_table(), pods, now)
        b2, r2 = oracle_lib.solve(recs, layout.speed_table(), pods, now, threads=threads)
        assert helpers.binding_bytes_equal(b1, b2), helpers.first_binding_diff(b1, b2)
        assert r1.tobytes() == r2.tobytes()


@pytest.mark.parametrize('config', [3, 5])
def test_mt_oracle_on_baseline_shapes(oracle_lib, config):
    import workload
    recs, speed, pods, now = workload.make_workload(config, n_nodes=1500, n_pods=400)
    b1, r1 = oracle_lib.solve(recs, speed, pods, now)
    b2, r2 = oracle_lib.solve(recs, speed, pods, now, threads=7)
    assert helpers.binding_bytes_equal(b1, b2), helpers.first_binding_diff(b1, b2)
    assert r1.tobytes() == r2.tobytes()
    assert (b1['status'] == 0).sum() > 100
