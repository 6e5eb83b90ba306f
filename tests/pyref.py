"""Tiny Python restatements used as ground truth by tests: they run the reference's own
expressions on REAL CPython sets, so any set-order model can be checked against the
interpreter that is running the test."""
import itertools


def choose_mapping(K, G, maskA, maskB, maskC):
    """Matcher.py:113-141,175-220,344-368,428-444 on real sets.  masks are sets of tuple indices
    (product order).  Returns (gtuple index, misc numa) or None."""
    prodG = list(itertools.product(range(K), repeat=G))
    prodB = list(itertools.product(range(K), repeat=G + 1))
    stmp = set()
    for i, p in enumerate(prodG):
        if i in maskA:
            stmp.add(p)
    a_list = list(stmp)
    stmp = set()
    for i, p in enumerate(prodB):
        if i in maskB:
            stmp.add(p)
    b_list = list(stmp)
    nic_tuples = [p for i, p in enumerate(prodG) if i in maskC]
    if not a_list or not b_list or not nic_tuples:
        return None
    gpu_tuples = [x for x in a_list]
    cpu_tuples = [x[:-1] for x in b_list]
    intersect = list(set(gpu_tuples) & set(cpu_tuples) & set(nic_tuples))
    if len(intersect) == 0:
        return None
    gl = a_list
    diff = set(a_list) - set(intersect)
    if len(diff):
        gl = intersect

    def node_delta(x):
        el = [gl[x].count(y) for y in range(K)]
        return max(el) - min(el)
    gidx, gval = 0, node_delta(0)
    for t in range(1, len(gl)):
        tmp = node_delta(t)
        if tmp > gval:
            gidx, gval = t, tmp
    gtuple = gl[gidx]
    cabbr = [x[:-1] for x in b_list]
    ctuple = b_list[cabbr.index(gtuple)]
    return prodG.index(gtuple), ctuple[-1]


# ----------------------------------------------------------------------------------------------------
# The reference's own Python path on packed workloads (bench.py's optional `python_reference` baseline and
# tools/python_reference_baseline.py).  Needs the unmodified reference (oracle/ref_loader): build container only.
# ----------------------------------------------------------------------------------------------------
def nodes_from_records(recs, speed_table, node_mod, group_names=None):
    """Packed nhd_node_rec records -> Node objects of `node_mod` (the reference's nhd.Node or the mirror) through
    their own ParseLabels (NFD label dictionaries), then the occupancy the records carry."""
    nodes = {}
    for i, r in enumerate(recs):
        nd = node_def_from_record(i, r, speed_table, group_names)
        n = node_mod.Node(nd['name'], nd['active'])
        assert n.ParseLabels(nd['labels'])
        n.SetHugepages(nd['hp_alloc'], nd['hp_free'])
        apply_occupancy(n, r)
        nodes[nd['name']] = n
    return nodes


def apply_occupancy(n, r):
    """What the record holds beyond the labels: cores / GPUs / NICs in use, the busy stamp."""
    used = [int(x) for x in r['used']]
    for c, core in enumerate(n.cores):
        if (used[c >> 6] >> (c & 63)) & 1:
            core.used = True
    for g, gpu in enumerate(n.gpus):
        gpu.used = bool((int(r['gpu_used']) >> g) & 1)
    for j, nic in enumerate(n.nics):
        if (int(r['nic_inuse']) >> j) & 1:
            nic.pods_used = 1
    n.busy_time = float(r['busy_time'])


def node_def_from_record(i, r, speed_table, group_names=None):
    """One packed record -> the node description tests/scenarios.make_node builds (name, NFD labels, hugepages)."""
    from tests import scenarios
    if True:
        K, phys = int(r['n_numa']), int(r['phys_cores'])
        smt = bool(int(r['flags']) & 1)
        gpus = [(g, next(k for k in range(K) if (int(r['gpu_numa_mask'][k]) >> g) & 1), (int(r['gpu_sw']) >> (4 * g)) & 0xF)
                for g in range(int(r['n_gpus']))]
        nics = []
        for j in range(int(r['n_nics'])):
            numa = next(k for k in range(K) if (int(r['nic_numa_mask'][k]) >> j) & 1)
            sw = (int(r['nic_sw'][j >> 4]) >> (4 * (j & 15))) & 0xF
            sc = (int(r['nic_speed'][j >> 4]) >> (4 * (j & 15))) & 0xF
            nics.append((f'eth{j}', int(round(float(speed_table[sc]) * 1000)), numa, sw))
        gm = int(r['group_mask'])
        groups = [(group_names[b] if group_names else ('default' if b == 0 else f'g{b}')) for b in range(64) if (gm >> b) & 1]
        return scenarios.make_node(f'n{i:06d}', sockets=K, phys_cores=phys, smt=smt, reserved_per_socket=2, gpus=gpus, nics=nics,
                                   groups=groups, hp_alloc=max(1, int(r['free_hugepages_gb'])), hp_free=int(r['free_hugepages_gb']),
                                   active=bool(int(r['flags']) & 2), maintenance=bool(int(r['flags']) & 4))


def pod_dict_from_record(p):
    """nhd_pod record -> the scenario dictionary tests/scenarios.build_top() turns into a CfgTopology with the same
    request vectors (CfgTopology.py:199-232): RX / TX sums on one core each, the rest plain workers."""
    from nhd_b200 import wire
    groups = []
    for g in range(int(p['n_groups'])):
        pg = p['groups'][g]
        n_proc = int(pg['n_proc'])
        fl = int(pg['flags'])
        pairs = [(float(pg['rx_gbps']), float(pg['tx_gbps']))] if fl & wire.GRP_HAS_NIC_CORES else []
        groups.append({'pairs': [list(x) for x in pairs], 'workers': n_proc - 2 * len(pairs),
                       'gpus': [int(pg['gpu_feeders'][j]) for j in range(int(pg['n_gpus']))],
                       'helpers': int(pg['n_helpers']), 'proc_smt': bool(fl & wire.GRP_PROC_SMT),
                       'helper_smt': bool(fl & wire.GRP_HELPER_SMT)})
    gm = int(p['group_mask'])
    return {'groups': [('default' if b == 0 else f'g{b}') for b in range(64) if (gm >> b) & 1],
            'map_type': {1: 'NUMA', 2: 'PCI'}.get(int(p['map_type']), 'NONE'), 'misc': int(p['n_misc']),
            'misc_smt': bool(int(p['flags']) & 1), 'hugepages': int(p['hugepages_gb']), 'proc_groups': groups}


def time_reference(recs, speed_table, pods, now, check_against=None):
    """Schedules `pods` on `recs` with the UNMODIFIED reference (one Python thread, logging disabled), as
    NHDScheduler.AttemptScheduling would; returns decisions/s and, when bindings are given, whether node and
    mapping agree."""
    import contextlib
    import io
    import os
    import time
    from oracle import ref_loader
    from tests import scenarios
    ref = ref_loader.load()
    t0 = time.perf_counter()
    nodes = nodes_from_records(recs, speed_table, ref.node)
    build_s = time.perf_counter() - t0
    names = list(nodes)
    matcher = ref.matcher.Matcher()
    agree = True
    t0 = time.perf_counter()
    for i, (p, t) in enumerate(zip(pods, now)):
        pd = pod_dict_from_record(p)
        top = scenarios.build_top(pd, ref.cfg)
        with contextlib.redirect_stdout(io.StringIO()):
            r = ref_loader.attempt_scheduling(ref, matcher, nodes, top, pd['groups'], float(t))
        if check_against is not None:
            want = int(check_against[i]['node'])
            got = names.index(r['node']) if r['status'] != 'none' else -1
            agree = agree and want == got
    dt = time.perf_counter() - t0
    return {'value': len(pods) / dt, 'unit': 'decisions/s', 'cores': 1, 'kind': 'reference (unmodified Python path)',
            'sample': f'first {len(pods)} pods on all {len(recs)} nodes, one thread, logging disabled ({dt:.1f} s; '
                      f'{build_s:.1f} s to build the Node objects from NFD labels)',
            'nodes_agree_with_bindings': bool(agree) if check_against is not None else None,
            'host_cores_available': os.cpu_count()}
