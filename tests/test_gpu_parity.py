"""GPU parity tests proper: the CUDA solver, called through the C-ABI, must produce the same
bindings and the same final node records as the oracle, bit for bit."""
import numpy as np
import pytest

import workload
from tests import helpers, ref_compare, scenarios

pytestmark = pytest.mark.gpu

# constant-clock batches use standing decisions per pod type and, when no CPU-only pod can spill, sweep the two pod
# classes side by side on two CTAs: every way of running the sweep must agree with the rest byte for byte
ALL_CPU_WARPS = (1, 2)          # 1: never side by side; 2: default


@pytest.fixture(scope='module')
def solver_mod():
    from nhd_b200 import solver
    return solver


def _run_cuda(solver_mod, recs, speed, pods, now, min_busy=30.0, extra_cpu_warps=()):
    """Runs the batch with the default sweep (speculating CPU-class warps + one GPU-class warp when the
    clock is constant), with the forced one-warp sweep and with one / two CPU-class warps; all must agree
    byte for byte."""
    outs = []
    modes = ((False, 0, 0), (True, 0, 0)) + tuple((False, c, 0) for c in extra_cpu_warps)
    if extra_cpu_warps:
        modes += ((False, 0, 1),               # constant-clock sweep without standing decisions (general path, table-driven evaluation)
                  (False, 0, 5))               # ... and without the direct-path tables (warp-wide evaluation)
    for single, cw, dbg in modes:
        s = solver_mod.Solver(speed, min_busy_secs=min_busy, single_warp=single, cpu_warps=cw, sweep_debug=dbg)
        try:
            s.load_nodes(recs)
            b = s.solve_batch(pods, now)
            final = s.read_nodes()
            outs.append((b, final, s.timing()))
        except Exception as e:
            raise AssertionError(f'sweep mode single_warp={single} cpu_warps={cw} debug={dbg}: {e}') from e
        finally:
            s.close()
    for k, o in enumerate(outs[1:]):
        assert helpers.binding_bytes_equal(outs[0][0], o[0]), (k + 1, helpers.first_binding_diff(outs[0][0], o[0]))
        assert outs[0][1].tobytes() == o[1].tobytes(), k + 1
    return outs[0]


@pytest.mark.parametrize('flavor', ['mixed', 'wild', 'vf', 'big'])
def test_random_scenarios_match_oracle(oracle_lib, solver_mod, flavor):
    placed = 0
    for seed in range(25):
        scn = scenarios.random_scenario(5000 + seed * 17 + len(flavor), n_nodes=10, n_pods=48, flavor=flavor,
                                        max_groups=4 if flavor != 'wild' else 3)
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        ob, orecs = oracle_lib.solve(recs, layout.speed_table(), pods, now)
        cb, crecs, _ = _run_cuda(solver_mod, recs, layout.speed_table(), pods, now)
        assert helpers.binding_bytes_equal(ob, cb), (seed, helpers.first_binding_diff(ob, cb))
        assert orecs.tobytes() == crecs.tobytes(), (seed, ref_compare.diff_records(orecs, crecs)[:3])
        placed += int((ob['status'] == 0).sum())
    assert placed > 100


def test_objects_near_the_packed_limits_match_oracle(oracle_lib, solver_mod):
    """Nodes with up to 256 logical cores, 16 GPUs, 32 NICs and 1 / 2 / 4 NUMA nodes; pods with up to 3 groups and
    72 cores (include/nhd_b200.h NHD_MAX_*), constant and moving clocks, every sweep mode."""
    placed = 0
    for seed in range(24):
        scn = scenarios.huge_scenario(81000 + seed, min_busy_secs=30.0 if seed % 3 else 0.0)
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        ob, orecs = oracle_lib.solve(recs, layout.speed_table(), pods, now, min_busy_secs=scn['min_busy_secs'])
        cb, crecs, _ = _run_cuda(solver_mod, recs, layout.speed_table(), pods, now, min_busy=scn['min_busy_secs'],
                                 extra_cpu_warps=(2, 5) if seed % 2 else ())
        assert helpers.binding_bytes_equal(ob, cb), (seed, helpers.first_binding_diff(ob, cb))
        assert orecs.tobytes() == crecs.tobytes(), (seed, ref_compare.diff_records(orecs, crecs)[:3])
        placed += int((ob['status'] == 0).sum())
    assert placed > 500


def test_filter_kernel_matches_oracle_candidates(oracle_lib, solver_mod):
    """F[type][node] from filter_kernel == filts[1] membership after IntersectResources
    (busy window aside), on partially filled clusters, plus the NOGPU / BUSY bitmaps."""
    for seed in range(6):
        scn = scenarios.random_scenario(900 + seed, n_nodes=300 + 37 * seed, n_pods=60, flavor='mixed')
        recs, pods, now, layout = ref_compare.pack_scenario(scn)
        _, filled = oracle_lib.solve(recs, layout.speed_table(), pods[:40], now[:40])
        s = solver_mod.Solver(layout.speed_table())
        try:
            s.load_nodes(filled)
            tail, tnow = pods[40:], np.full(20, 2000.0)
            s.stage_batch(tail, tnow)
            feas, nogpu, busy, pt = s.filter_bitmaps()
        finally:
            s.close()
        assert np.array_equal(nogpu, filled['n_gpus'] == 0)
        assert np.array_equal(busy, (2000.0 - filled['busy_time']) < 30.0)
        for i, pod in enumerate(tail):
            cand = oracle_lib.candidates(filled, layout.speed_table(), pod, now=1e9).astype(bool)
            # with several distinct pod group lists the node-group gate is applied by the sweep, not the filter
            elig = (filled['group_mask'] & pod['group_mask']) != 0
            assert np.array_equal(feas[pt[i]] & elig, cand), (seed, i)


@pytest.mark.parametrize('config,n_nodes,n_pods', [(1, None, None), (2, None, None), (3, 4096, 384), (5, 4096, 384)])
def test_baseline_configs_match_oracle(oracle_lib, solver_mod, config, n_nodes, n_pods):
    recs, speed, pods, now = workload.make_workload(config, n_nodes, n_pods)
    ob, orecs = oracle_lib.solve(recs, speed, pods, now)
    cb, crecs, timing = _run_cuda(solver_mod, recs, speed, pods, now, extra_cpu_warps=ALL_CPU_WARPS)
    assert helpers.binding_bytes_equal(ob, cb), helpers.first_binding_diff(ob, cb)
    assert orecs.tobytes() == crecs.tobytes()
    assert timing['n_launches'] == 6      # direct-path tables, filter, sweep, resolve, core ids, commit


@pytest.mark.parametrize('wild', [False, True], ids=['regular', 'heterogeneous'])
def test_many_pod_type_catalogues_match_oracle(oracle_lib, solver_mod, wild):
    """The table-driven direct path over many different pod-type catalogues (group counts, SMT flags, misc cores,
    NIC demands, hugepages), on clusters big enough for the no-spill certificate to hold (the two pod classes on two
    CTAs) — and on a heterogeneous cluster where most hardware classes fall outside the tables."""
    placed = 0
    for seed in range(10):
        recs, speed = workload.make_cluster(3 + seed % 2, n_nodes=900, seed=4200 + seed, wild=wild)
        types = workload.make_pod_types(3 + seed % 2, seed=9000 + seed)
        rng = np.random.default_rng(seed)
        # skew the stream so that some types fill their nodes quickly
        pick = rng.choice(len(types), size=220, p=rng.dirichlet(np.ones(len(types)) * 0.7))
        pods = types[pick].copy()
        now = np.full(len(pods), 777.0)
        ob, orecs = oracle_lib.solve(recs, speed, pods, now)
        cb, crecs, _ = _run_cuda(solver_mod, recs, speed, pods, now, extra_cpu_warps=ALL_CPU_WARPS)
        assert helpers.binding_bytes_equal(ob, cb), (seed, helpers.first_binding_diff(ob, cb))
        assert orecs.tobytes() == crecs.tobytes(), seed
        placed += int((ob['status'] == 0).sum())
    assert placed > 1500


def test_clock_changes_and_busy_window(oracle_lib, solver_mod):
    """Per-pod clocks that move forwards, stall and jump backwards exercise the busy list."""
    recs, speed, pods, _ = workload.make_workload(3, n_nodes=512, n_pods=200)
    rng = np.random.default_rng(3)
    now = 1000.0 + np.cumsum(rng.choice([0.0, 0.0, 5.0, 12.0, 31.0, -40.0], size=len(pods)))
    ob, orecs = oracle_lib.solve(recs, speed, pods, now)
    cb, crecs, _ = _run_cuda(solver_mod, recs, speed, pods, now)
    assert helpers.binding_bytes_equal(ob, cb), helpers.first_binding_diff(ob, cb)
    assert orecs.tobytes() == crecs.tobytes()


def test_update_snapshot_restore_and_prefix(oracle_lib, solver_mod):
    recs, speed, pods, now = workload.make_workload(3, n_nodes=2048, n_pods=256)
    s = solver_mod.Solver(speed)
    try:
        s.load_nodes(recs)
        s.snapshot()
        full = s.solve_batch(pods, now)
        s.restore()
        again = s.solve_batch(pods, now)
        assert helpers.binding_bytes_equal(full, again)                 # deterministic replay
        s.restore()
        half = s.solve_batch(pods[:128], now[:128])
        assert helpers.binding_bytes_equal(full[:128], half)           # pod i only sees pods < i
        rest = s.solve_batch(pods[128:], now[128:])                    # state carried across batches
        assert helpers.binding_bytes_equal(full[128:], rest)
        # release: put back the original record of every touched node, then re-solve
        s.restore()
        s.solve_batch(pods[:64], now[:64])
        touched = np.unique(full[:64]['node'][full[:64]['node'] >= 0])
        s.update_nodes(touched, recs[touched])
        assert s.read_nodes().tobytes() == recs.tobytes()
        assert helpers.binding_bytes_equal(s.solve_batch(pods, now), full)
    finally:
        s.close()


def test_full_size_properties(oracle_lib, solver_mod):
    """65 536 nodes x 4 096 pods (the metric's configuration): too big for the oracle in
    full, so (a) the first pods are compared with the oracle at full N, (b) size-independent
    invariants are checked on everything."""
    recs, speed, pods, now = workload.make_workload(4)
    cb, crecs, timing = _run_cuda(solver_mod, recs, speed, pods, now, extra_cpu_warps=ALL_CPU_WARPS)
    n_check = 12
    ob, _ = oracle_lib.solve(recs, speed, pods[:n_check], now[:n_check])
    assert helpers.binding_bytes_equal(ob, cb[:n_check]), helpers.first_binding_diff(ob, cb[:n_check])
    placed = cb[cb['status'] == 0]
    assert len(placed) > 3000
    # every core handed out was free before, and is handed out once
    used0 = np.unpackbits(recs['used'].view(np.uint8).reshape(len(recs), 32), axis=1, bitorder='little')
    used1 = np.unpackbits(crecs['used'].view(np.uint8).reshape(len(recs), 32), axis=1, bitorder='little')
    taken = np.zeros_like(used0)
    for b in placed:
        cores = b['cores'][:b['n_cores']]
        assert len(set(cores.tolist())) == len(cores)
        assert not used0[b['node'], cores].any()
        assert not taken[b['node'], cores].any()
        taken[b['node'], cores] = 1
    assert np.array_equal(used1, used0 | taken)
    # GPU pods: one per node per busy window (constant clock), on GPU nodes only
    gpu_pods = placed[placed['n_gpus'] > 0]
    assert len(np.unique(gpu_pods['node'])) == len(gpu_pods)
    assert (recs['n_gpus'][gpu_pods['node']] > 0).all()
    # untouched nodes are bit-identical
    untouched = np.ones(len(recs), bool)
    untouched[placed['node']] = False
    assert recs[untouched].tobytes() == crecs[untouched].tobytes()
    assert timing['n_types'] <= 16


def test_constant_clock_random_scenarios(oracle_lib, solver_mod):
    """Constant clock => the multi-warp sweep (GPU pods / CPU-only pods concurrently, CPU-only pods worked out
    ahead of their turn, spills serialised); small clusters make consecutive pods hit the same node and make
    CPU-only pods spill onto GPU nodes all the time."""
    spills = 0
    for seed in range(12):
        scn = scenarios.random_scenario(8000 + seed * 3, n_nodes=12, n_pods=64, flavor='mixed' if seed % 3 else 'big',
                                        max_groups=3)
        recs, pods, _, layout = ref_compare.pack_scenario(scn)
        for busy in (30.0, 0.0):
            now = np.full(len(pods), 5000.0)
            ob, orecs = oracle_lib.solve(recs, layout.speed_table(), pods, now, min_busy_secs=busy)
            cb, crecs, _ = _run_cuda(solver_mod, recs, layout.speed_table(), pods, now, min_busy=busy,
                                     extra_cpu_warps=ALL_CPU_WARPS)
            assert helpers.binding_bytes_equal(ob, cb), (seed, busy, helpers.first_binding_diff(ob, cb))
            assert orecs.tobytes() == crecs.tobytes()
            placed = ob[ob['status'] == 0]
            cpu_only = placed[placed['n_gpus'] == 0]
            spills += int((recs['n_gpus'][cpu_only['node']] > 0).sum())
    assert spills > 8


def test_upload_validation_and_pinned_buffers(oracle_lib, solver_mod):
    """Records are validated on the device during nhd_load_nodes; pinned buffers take the zero-copy path."""
    recs, speed, pods, now = workload.make_workload(3, n_nodes=1000, n_pods=64)
    s = solver_mod.Solver(speed)
    try:
        bad = recs.copy()
        bad['n_numa'][517] = 7
        with pytest.raises(solver_mod.SolverError) as ei:
            s.load_nodes(bad)
        assert ei.value.code == -2 and '517' in str(ei.value)
        with pytest.raises(solver_mod.SolverError):
            s.solve_batch(pods, now)                       # nothing loaded after a rejected upload
        prec = solver_mod.pinned_array(len(recs), recs.dtype)
        prec[:] = recs
        ppods = solver_mod.pinned_array(len(pods), pods.dtype)
        ppods[:] = pods
        out = solver_mod.pinned_array(len(pods), ref_compare.wire.BINDING_DTYPE)
        s.load_nodes(prec)
        s.solve_batch(ppods, now, out=out)
        ob, orecs = oracle_lib.solve(recs, speed, pods, now)
        assert helpers.binding_bytes_equal(ob, out)
        assert s.read_nodes().tobytes() == orecs.tobytes()
    finally:
        s.close()


def test_edge_cases_empty_and_rejected_inputs(oracle_lib, solver_mod):
    """Empty batches / clusters, invalid map types, unschedulable pods, descriptors outside the limits."""
    recs, speed, pods, now = workload.make_workload(3, n_nodes=300, n_pods=40)
    s = solver_mod.Solver(speed)
    try:
        s.load_nodes(recs)
        assert len(s.solve_batch(pods[:0], now[:0])) == 0                     # empty batch
        assert s.read_nodes().tobytes() == recs.tobytes()
        weird = pods.copy()
        weird['map_type'][::3] = 3                                            # TOPOLOGY_MAP_NONE -> (None,)  (Matcher.py:45-47)
        weird['hugepages_gb'][1::3] = 10_000                                  # nobody has that many hugepages
        weird['group_mask'][2::7] = 1 << 40                                   # a node group no node carries
        ob, orecs = oracle_lib.solve(recs, speed, weird, now)
        cb = s.solve_batch(weird, now)
        assert helpers.binding_bytes_equal(ob, cb), helpers.first_binding_diff(ob, cb)
        assert s.read_nodes().tobytes() == orecs.tobytes()
        assert set(np.unique(cb['status'])) >= {0, 1, 4}
        bad = pods[:1].copy()
        bad['n_groups'] = 0
        with pytest.raises(solver_mod.SolverError):
            s.solve_batch(bad, now[:1])
        bad = pods[:1].copy()
        bad['groups'][0][0]['n_proc'] = 200                                   # > NHD_MAX_POD_CORES
        with pytest.raises(solver_mod.SolverError) as ei:
            s.solve_batch(bad, now[:1])
        assert ei.value.code == -2
        s.load_nodes(recs[:0])                                                # empty cluster
        out = s.solve_batch(pods[:5], now[:5])
        assert (out['status'] == 1).all() and (out['node'] == -1).all()
        s.load_nodes(recs[:1])                                                # one node, many pods: it fills up
        ob, orecs = oracle_lib.solve(recs[:1], speed, pods, now)
        cb = s.solve_batch(pods, now)
        assert helpers.binding_bytes_equal(ob, cb) and s.read_nodes().tobytes() == orecs.tobytes()
    finally:
        s.close()


def test_tuple_limit_follows_the_schedulable_nodes(oracle_lib, solver_mod):
    """numa^(groups+1) <= NHD_MAX_TUPLES is judged against the nodes the filter could accept: a 4-socket node that is
    inactive (or in maintenance) does not make 4-group pods unsupported cluster-wide, and the limit follows updates in
    both directions (Matcher.py:73-75 never looks at such a node either)."""
    from nhd_b200 import wire
    def nics(sockets):
        return [(f'vf{i}', 100000, i % sockets, 0x10 * (i % sockets + 1)) for i in range(2 * sockets)]
    nodes = [scenarios.make_node('a', sockets=2, phys_cores=32, nics=nics(2)),
             scenarios.make_node('b', sockets=4, phys_cores=64, nics=nics(4)),
             scenarios.make_node('c', sockets=2, phys_cores=32, nics=nics(2))]
    pod4 = scenarios.make_pod([scenarios.make_group(pairs=((5, 5),), workers=1) for _ in range(4)], misc=1)
    pod2 = scenarios.make_pod([scenarios.make_group(pairs=((5, 5),), workers=2) for _ in range(2)], misc=1)
    scn = {'nodes': nodes, 'pods': [pod4, pod2, pod4, pod2], 'now': [1000.0] * 4, 'min_busy_secs': 0.0}
    recs, pods, now, layout = ref_compare.pack_scenario(scn)
    speed = layout.speed_table()
    off = recs.copy()
    off['flags'][1] &= 0xFF ^ wire.NODE_ACTIVE
    maint = recs.copy()
    maint['flags'][1] |= wire.NODE_MAINTENANCE
    s = solver_mod.Solver(speed, min_busy_secs=0.0)
    try:
        for variant in (off, maint):
            s.load_nodes(variant)
            ob, orecs = oracle_lib.solve(variant, speed, pods, now, min_busy_secs=0.0)
            cb = s.solve_batch(pods, now)
            assert helpers.binding_bytes_equal(ob, cb), helpers.first_binding_diff(ob, cb)
            assert s.read_nodes().tobytes() == orecs.tobytes()
            assert (cb['status'] == 0).all() and (cb['node'] != 1).all()
        s.load_nodes(off)
        s.update_nodes(np.array([1], dtype=np.int32), recs[1:2])              # the 4-socket node comes back
        with pytest.raises(solver_mod.SolverError) as ei:
            s.solve_batch(pods, now)
        assert ei.value.code == -2
        two = s.solve_batch(pods[1::2], now[1::2])                            # 2-group pods: 4^3 = 64 tuples, fine
        ob, _ = oracle_lib.solve(recs, speed, pods[1::2], now[1::2], min_busy_secs=0.0)
        assert helpers.binding_bytes_equal(ob, two)
        s.load_nodes(recs)
        s.update_nodes(np.array([1], dtype=np.int32), off[1:2])               # ... and leaves again: the limit shrinks
        ob, orecs = oracle_lib.solve(off, speed, pods, now, min_busy_secs=0.0)
        cb = s.solve_batch(pods, now)
        assert helpers.binding_bytes_equal(ob, cb) and s.read_nodes().tobytes() == orecs.tobytes()
    finally:
        s.close()
