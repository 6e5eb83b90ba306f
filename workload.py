"""Synthetic clusters and pending-pod streams for the five BASELINE.json configs
(SURVEY.md section 8d), generated directly as packed wire records with numpy so that the
64k / 256k-node cases take well under a second to build.  Benchmark and test tooling, not
part of the solver.

    config 1:      16 nodes x     4 pods   2-socket STD nodes (the oracle KAT shape)
    config 2:   4 096 nodes x   256 pods   CPU-core + NUMA predicates only
    config 3:  16 384 nodes x 1 024 pods   + GPU / NIC PCI-switch affinity (50 % GPU pod types)
    config 4:  65 536 nodes x 4 096 pods   same mix (the metric's configuration)
    config 5: 262 144 nodes x 8 192 pods   + SR-IOV VFs (8 x 25G per NUMA) + node groups N:N

Every node: K = 2 sockets; 32/48/64 physical cores (p = .25/.5/.25); SMT with p = .75; the
first two physical cores of each socket (and their siblings) reserved for the OS; hugepages
32/64/128 GB; half the nodes carry 8 GPUs (4 per NUMA node, 2 per PCIe switch); configs 1-4
have one or two 100G NICs per NUMA node on that NUMA node's switches.  30 % of the nodes
start partially occupied.  Pod streams draw uniformly from 16 pod types; `now` is constant.
"""
import numpy as np

from nhd_b200 import wire

CONFIGS = {1: (16, 4), 2: (4096, 256), 3: (16384, 1024), 4: (65536, 4096), 5: (262144, 8192)}
SEED0 = 20260922
NOW0 = 1000.0
N_TYPES = 16
N_GROUP_NAMES = 16


def _range_mask(lo, hi):
    """uint64[.., 4] mask of bits [lo, hi) for arrays lo, hi (0..256)."""
    lo = np.asarray(lo, dtype=np.int64)[..., None]
    hi = np.asarray(hi, dtype=np.int64)[..., None]
    base = np.arange(4, dtype=np.int64) * 64
    a = np.clip(lo - base, 0, 64)
    b = np.clip(hi - base, 0, 64)

    def low_bits(n):      # (1 << n) - 1 for n in 0..64, as uint64
        n = n.astype(np.uint64)
        full = n >= 64
        v = (np.uint64(1) << np.where(full, np.uint64(0), n)) - np.uint64(1)
        return np.where(full, np.uint64(0xFFFFFFFFFFFFFFFF), v)
    return np.where(b > a, low_bits(b) & ~low_bits(a), np.uint64(0))


def make_cluster(config: int, n_nodes: int = None, seed: int = None, wild: bool = False):
    """Returns (records[NODE_DTYPE], speed_table[16]) for a BASELINE config.
    ``wild``: a heterogeneous variant of configs 1-4 (sensitivity runs, not a BASELINE config): one to four NICs per
    NUMA node at 25 / 40 / 100 G in any mix — hundreds of hardware classes and more NIC signatures than the sweep's
    direct-path tables hold, so a large share of the decisions takes the general path."""
    N = CONFIGS[config][0] if n_nodes is None else n_nodes
    rng = np.random.default_rng(SEED0 + config if seed is None else seed)
    recs = np.zeros(N, dtype=wire.NODE_DTYPE)
    phys = rng.choice([32, 48, 64], size=N, p=[.25, .5, .25]).astype(np.int64)
    smt = rng.random(N) < 0.75
    per = phys // 2
    recs['n_numa'] = 2
    recs['phys_cores'] = phys
    recs['flags'] = wire.NODE_ACTIVE | np.where(smt, wire.NODE_SMT, 0)
    hp = rng.choice([32, 64, 128], size=N)
    recs['free_hugepages_gb'] = hp
    recs['busy_time'] = 0.0

    # reserved OS cores: first two physical cores of each socket (+ SMT siblings)
    used = _range_mask(0, 2) | _range_mask(per, per + 2)
    used_s = _range_mask(phys, phys + 2) | _range_mask(phys + per, phys + per + 2)
    used = np.where(smt[:, None], used | used_s, used)

    # GPUs: half of the nodes, 8 GPUs: gpu i -> numa i // 4, switch i // 2
    has_gpu = rng.random(N) < 0.5
    recs['n_gpus'] = np.where(has_gpu, 8, 0)
    recs['gpu_numa_mask'][:, 0] = np.where(has_gpu, 0x0F, 0)
    recs['gpu_numa_mask'][:, 1] = np.where(has_gpu, 0xF0, 0)
    gpu_sw = sum((i // 2) << (4 * i) for i in range(8))
    recs['gpu_sw'] = np.where(has_gpu, np.uint64(gpu_sw), np.uint64(0))

    # NICs
    if config == 5:
        # one SR-IOV PF per NUMA node (excluded, Node.py:380-395) + 8 VFs per NUMA node at 25G;
        # VF j of NUMA k sits on switch 2k + (j % 2)
        speeds = [25.0]
        recs['n_nics'] = 16
        recs['nic_numa_mask'][:, 0] = 0x00FF
        recs['nic_numa_mask'][:, 1] = 0xFF00
        sw = sum((2 * (i // 8) + (i % 2)) << (4 * i) for i in range(16))
        recs['nic_sw'][:, 0] = np.uint64(sw)
        recs['nic_speed'][:] = 0
        n_names = rng.integers(1, 4, size=N)
        gm = np.zeros(N, dtype=np.uint64)
        for j in range(3):
            pick = rng.integers(0, N_GROUP_NAMES, size=N).astype(np.uint64)
            gm |= np.where(j < n_names, np.uint64(1) << pick, np.uint64(0))
        recs['group_mask'] = gm
    elif wild:
        speeds = [100.0, 40.0, 25.0]
        n0 = rng.integers(1, 5, size=N)
        n1 = rng.integers(1, 5, size=N)
        recs['n_nics'] = n0 + n1
        recs['nic_numa_mask'][:, 0] = (1 << n0) - 1
        recs['nic_numa_mask'][:, 1] = ((1 << n1) - 1) << n0
        sw = np.zeros(N, dtype=np.uint64)
        sp = np.zeros(N, dtype=np.uint64)
        for j in range(8):
            on1 = j >= n0                                     # NIC j sits on NUMA 1 (list order numa0.., numa1..)
            live = j < n0 + n1
            swj = np.where(on1, 2 + ((j - n0) % 2), j % 2).astype(np.uint64)
            sw |= np.where(live, swj << np.uint64(4 * j), np.uint64(0))
            sp |= np.where(live, rng.integers(0, 3, size=N).astype(np.uint64) << np.uint64(4 * j), np.uint64(0))
        recs['nic_sw'][:, 0] = sw
        recs['nic_speed'][:, 0] = sp
        recs['group_mask'] = 1
    else:
        # one or two 100G NICs per NUMA node, list order numa0.., numa1..; NIC j of NUMA k on switch 2k + j
        speeds = [100.0]
        n0 = rng.integers(1, 3, size=N)
        n1 = rng.integers(1, 3, size=N)
        recs['n_nics'] = n0 + n1
        recs['nic_numa_mask'][:, 0] = (1 << n0) - 1
        recs['nic_numa_mask'][:, 1] = ((1 << n1) - 1) << n0
        sw = np.zeros(N, dtype=np.uint64)
        for j in range(2):
            sw |= np.where(j < n0, np.uint64(j) << np.uint64(4 * j), np.uint64(0))
            sw |= np.where(j < n1, np.uint64(2 + j) << (np.uint64(4) * (n0 + j).astype(np.uint64)), np.uint64(0))
        recs['nic_sw'][:, 0] = sw
        recs['group_mask'] = 1          # everybody in 'default' (bit 0)

    # pre-occupancy: 30 % of the nodes already run something
    pre = rng.random(N) < 0.30
    u0 = np.where(pre, rng.integers(0, per - 1), 0)          # busy physical cores per socket, lowest ids first
    u1 = np.where(pre, rng.integers(0, per - 1), 0)
    occ = _range_mask(2, np.minimum(2 + u0, per)) | _range_mask(per + 2, np.minimum(per + 2 + u1, 2 * per))
    occ_s = _range_mask(phys + 2, np.minimum(phys + 2 + u0, phys + per)) | \
        _range_mask(phys + per + 2, np.minimum(phys + per + 2 + u1, phys + 2 * per))
    used = used | occ | np.where(smt[:, None], occ_s, np.uint64(0))
    recs['used'] = used
    g0 = np.where(pre & has_gpu, rng.integers(0, 5, size=N), 0)
    g1 = np.where(pre & has_gpu, rng.integers(0, 5, size=N), 0)
    recs['gpu_used'] = ((1 << g0) - 1) | (((1 << g1) - 1) << 4)
    nic_busy = np.where(pre, rng.integers(0, 4, size=N), 0)  # bit 0: first NIC of NUMA 0, bit 1: first of NUMA 1
    first1 = np.log2(np.maximum(recs['nic_numa_mask'][:, 1], 1) & -np.maximum(recs['nic_numa_mask'][:, 1], 1).astype(np.int64)).astype(np.int64)
    recs['nic_inuse'] = (nic_busy & 1) | (((nic_busy >> 1) & 1) << first1)
    recs['free_hugepages_gb'] = hp - np.where(pre, rng.integers(0, 16, size=N), 0)

    speed_table = np.zeros(wire.MAX_SPEED_CLASSES, dtype='<f8')
    speed_table[:len(speeds)] = speeds
    return recs, speed_table


def make_pod_types(config: int, seed: int = None):
    """The catalogue of 16 pod types of a config, as nhd_pod records (group_mask = default)."""
    rng = np.random.default_rng((SEED0 + 1000 + config) if seed is None else seed)
    types = np.zeros(N_TYPES, dtype=wire.POD_DTYPE)
    for t in range(N_TYPES):
        gpu_type = config >= 3 and t % 2 == 1
        p = types[t]
        G = int(rng.integers(1, 3))
        p['n_groups'] = G
        p['map_type'] = wire.MAP_PCI if gpu_type else wire.MAP_NUMA
        p['n_misc'] = int(rng.integers(1, 3))
        p['flags'] = wire.POD_MISC_SMT if rng.random() < 0.5 else 0
        p['hugepages_gb'] = int(rng.choice([0, 2, 4]))
        p['group_mask'] = 1
        for g in range(G):
            pg = p['groups'][g]
            n_proc = int(rng.integers(2, 7))
            has_pair = rng.random() < 0.7
            pg['n_proc'] = n_proc
            pg['n_helpers'] = int(rng.integers(0, 3))
            fl = (wire.GRP_PROC_SMT if rng.random() < 0.5 else 0) | (wire.GRP_HELPER_SMT if rng.random() < 0.5 else 0)
            if has_pair:
                pg['rx_gbps'] = float(rng.choice([0, 10, 20]))
                pg['tx_gbps'] = float(rng.choice([0, 10, 20]))
                fl |= wire.GRP_HAS_NIC_CORES
            if gpu_type:
                pg['n_gpus'] = 1
                pg['gpu_feeders'][0] = int(rng.integers(1, 3))
            pg['flags'] = fl
    return types


def make_pods(config: int, n_pods: int = None, seed: int = None):
    """Returns (pods[POD_DTYPE], now[f8]) — the pending-pod stream of a config."""
    P = CONFIGS[config][1] if n_pods is None else n_pods
    rng = np.random.default_rng((SEED0 + 2000 + config) if seed is None else seed)
    types = make_pod_types(config)
    pick = rng.integers(0, N_TYPES, size=P)
    pods = types[pick].copy()
    if config == 5:
        # per-pod node-group list: 1-2 of the 16 names (NHDScheduler.py:235-247)
        a = rng.integers(0, N_GROUP_NAMES, size=P).astype(np.uint64)
        b = rng.integers(0, N_GROUP_NAMES, size=P).astype(np.uint64)
        two = rng.random(P) < 0.5
        pods['group_mask'] = (np.uint64(1) << a) | np.where(two, np.uint64(1) << b, np.uint64(0))
    now = np.full(P, NOW0, dtype='<f8')
    return pods, now


def make_workload(config: int, n_nodes: int = None, n_pods: int = None, wild: bool = False):
    recs, speed = make_cluster(config, n_nodes, wild=wild)
    pods, now = make_pods(config, n_pods)
    return recs, speed, pods, now
