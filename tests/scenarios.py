"""Synthetic clusters and pod streams as plain data (JSON-able), plus builders that turn
them into reference-shaped objects using EITHER the reference's own classes
(nhd.Node / nhd.CfgTopology, build container only) or this repo's mirrors.

A scenario is
    {'nodes': [{'name', 'labels', 'hp_alloc', 'hp_free', 'active'}...],
     'pods':  [{'groups', 'map_type', 'misc', 'misc_smt', 'hugepages',
                'proc_groups': [{'pairs': [[rx, tx]...], 'workers', 'gpus': [feeders...],
                                 'helpers', 'proc_smt', 'helper_smt'}...]}...],
     'now':   [float per pod],
     'min_busy_secs': float}
Node labels use the NFD formats ``Node.ParseLabels`` reads (nhd/Node.py:327-454); pods
are built the way ``TriadCfgParser.ParseModGroups`` does (nhd/TriadCfgParser.py:134-304).
"""
import zlib

import numpy as np

NFD = 'feature.node.kubernetes.io/'
GROUP_NAMES = ['default'] + [f'grp{i}' for i in range(15)]


# ----------------------------------------------------------------------------------
# node label generation
# ----------------------------------------------------------------------------------
def _ranges(ids):
    """sorted ids -> "a-b,c-d" cpuset string"""
    ids = sorted(ids)
    out, i = [], 0
    while i < len(ids):
        j = i
        while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1:
            j += 1
        out.append(f'{ids[i]}-{ids[j]}' if j > i else f'{ids[i]}')
        i = j + 1
    return ','.join(out)


def make_node(name, sockets=2, phys_cores=32, smt=True, reserved_per_socket=2, gpus=(), nics=(),
              sriov_pfs=(), groups=None, hp_alloc=64, hp_free=None, active=True, maintenance=False,
              isol=True, vlan=100, res_hugepages=None, isolcpus=None):
    """gpus: [(device_id, numa, pciesw)], nics: [(ifname, mbps, numa, pciesw)] in label order."""
    labels = {NFD + 'nfd-extras-cpu.numSockets': str(sockets),
              NFD + 'nfd-extras-cpu.num_cores': str(phys_cores)}
    if smt:
        labels[NFD + 'cpu-hardware_multithreading'] = 'true'
    if isolcpus is not None:
        labels[NFD + 'nfd-extras-cpu.isolcpus'] = isolcpus
    elif isol:
        per = phys_cores // sockets
        iso = []
        for s in range(sockets):
            iso += list(range(s * per + reserved_per_socket, (s + 1) * per))
        if smt:
            iso += [c + phys_cores for c in iso]
        # the reference splits on '_' then parses "a-b,c" (Node.py:356-359)
        labels[NFD + 'nfd-extras-cpu.isolcpus'] = _ranges(iso).replace(',', '_')
    for dev, numa, sw in gpus:
        labels[NFD + f'nfd-extras-gpu.{dev}.V100.{numa}.{sw:x}'] = 'true'
    for pf in sriov_pfs:
        labels[NFD + f'nfd-extras-sriov.x.{pf}'] = 'true'
    for i, (ifname, mbps, numa, sw) in enumerate(nics):
        mac = f'{0x0c42a1000000 + (zlib.crc32(name.encode()) % 65536) * 256 + i:012x}'
        speed = f'{mbps}Mbs' if mbps else 'down'
        labels[NFD + f'nfd-extras-nic.{ifname}.mlx.{mac}.{speed}.{numa}.{sw:x}.{i + 1:x}.0'] = 'true'
    labels['DATA_PLANE_VLAN'] = str(vlan)
    labels['DATA_DEFAULT_GW'] = '10.1.0.1/24'
    if groups is not None:
        labels['NHD_GROUP'] = '.'.join(groups)
    if maintenance:
        labels['sigproc.viasat.io/maintenance'] = 'draining'
    if res_hugepages is not None:
        labels['RES_HUGEPAGES_GB'] = str(res_hugepages)
    return {'name': name, 'labels': labels, 'hp_alloc': hp_alloc,
            'hp_free': hp_alloc if hp_free is None else hp_free, 'active': active}


def make_pod(proc_groups, misc=1, misc_smt=False, map_type='NUMA', hugepages=0, groups=('default',)):
    return {'groups': list(groups), 'map_type': map_type, 'misc': misc, 'misc_smt': misc_smt,
            'hugepages': hugepages, 'proc_groups': proc_groups}


def make_group(pairs=((10, 10),), workers=1, gpus=(), helpers=0, proc_smt=False, helper_smt=False):
    return {'pairs': [list(p) for p in pairs], 'workers': workers, 'gpus': list(gpus),
            'helpers': helpers, 'proc_smt': proc_smt, 'helper_smt': helper_smt}


# ----------------------------------------------------------------------------------
# object builders (cfg_mod / node_mod are the reference's modules or this repo's mirrors)
# ----------------------------------------------------------------------------------
def build_nodes(scn, node_mod):
    nodes = {}
    for nd in scn['nodes']:
        n = node_mod.Node(nd['name'], nd.get('active', True))
        if not n.ParseLabels(nd['labels']):
            raise ValueError(f'labels of {nd["name"]} rejected')
        n.SetHugepages(nd['hp_alloc'], nd['hp_free'])
        nodes[nd['name']] = n
    return nodes


def build_top(pod, cfg_mod):
    C = cfg_mod
    top = C.CfgTopology()
    top.SetTopMapType(pod['map_type'])
    top.SetMiscCoreSmt(C.SMTSetting.SMT_ENABLED if pod['misc_smt'] else C.SMTSetting.SMT_DISABLED)
    top.SetCtrlVlan(C.VLANInfo('ctrl_vlan', 0))
    top.hugepages_gb = pod['hugepages']
    none, rx, tx = (C.NICCoreDirection.NIC_CORE_DIRECTION_NONE, C.NICCoreDirection.NIC_CORE_DIRECTION_RX,
                    C.NICCoreDirection.NIC_CORE_DIRECTION_TX)
    grp = C.NUMASetting.LOGICAL_NUMA_GROUP
    for i in range(pod['misc']):
        top.AddMiscCore(C.Core(f'ext{i}', 0, none, C.NUMASetting.LOGICAL_NUMA_DONT_CARE, -1))
    for gi, g in enumerate(pod['proc_groups']):
        pg = C.ProcGroup()
        pg.SetHelperSmt(C.SMTSetting.SMT_ENABLED if g['helper_smt'] else C.SMTSetting.SMT_DISABLED)
        pg.SetProcSmt(C.SMTSetting.SMT_ENABLED if g['proc_smt'] else C.SMTSetting.SMT_DISABLED)
        for h in range(g['helpers']):
            pg.AddMiscCore(C.Core(f'm{gi}.h{h}', 0, none, grp, -1))
        pg.SetDataVlan(C.VLANInfo(f'm{gi}.vlan', 0))
        for pi, (r, t) in enumerate(g['pairs']):
            rxc = C.Core(f'm{gi}.rx{pi}', r, rx, grp, -1)
            pg.AddGroupCore(rxc)
            txc = C.Core(f'm{gi}.tx{pi}', t, tx, grp, -1)
            pg.AddGroupCore(txc)
            top.AddNicPairing(rxc, txc)
        for w in range(g['workers']):
            pg.AddGroupCore(C.Core(f'm{gi}.w{w}', 0, none, grp, -1))
        for di, feeders in enumerate(g['gpus']):
            clist = [C.Core(f'm{gi}.g{di}.c{k}', 0, none, grp, -1) for k in range(feeders)]
            pg.AddGroupGPU(C.GPU(clist, [f'm{gi}.g{di}.dev'] * feeders, C.GpuType.GPU_TYPE_ALL, -1))
        top.proc_groups.append(pg)
    return top


def extract_result(top):
    """Everything SetPhysicalIdsFromMapping wrote into the topology, canonical order."""
    cores, gpus, macs, vlans = [], [], [], []
    for pg in top.proc_groups:
        for g in pg.group_gpus:
            gpus.append(g.device_id)
            cores += [c.core for c in g.cpu_cores]
        cores += [c.core for c in pg.proc_cores]
        cores += [c.core for c in pg.misc_cores]
        vlans.append(pg.vlan.vlan)
    cores += [c.core for c in top.misc_cores]
    macs = [p.mac for p in top.nic_core_pairing]
    return {'cores': cores, 'gpus': gpus, 'macs': macs, 'vlans': vlans,
            'ctrl_vlan': top.ctrl_vlan.vlan, 'gw': top.data_default_gw}


def node_state(node):
    """Mutable state of a Node object as plain data."""
    return {'used': [i for i, c in enumerate(node.cores) if c.used],
            'gpu_used': [i for i, g in enumerate(node.gpus) if g.used],
            'pods_used': [n.pods_used for n in node.nics],
            'speed_used': [[float(n.speed_used[0]), float(n.speed_used[1])] for n in node.nics],
            'free_hugepages': node.mem.free_hugepages_gb,
            'busy_time': node.busy_time}


# ----------------------------------------------------------------------------------
# random scenarios
# ----------------------------------------------------------------------------------
def random_node(rng, name, flavor='mixed', n_group_names=4):
    sockets = 2
    if flavor == 'wild':
        sockets = int(rng.choice([1, 2, 2, 2, 3, 4]))
    per = int(rng.choice([4, 6, 8, 12, 16])) if flavor != 'big' else int(rng.choice([16, 24, 32]))
    phys = per * sockets
    smt = bool(rng.random() < 0.7)
    reserved = int(rng.integers(0, min(3, per)))
    has_gpu = rng.random() < 0.55
    gpus, nics = [], []
    sw_base = [0x10 * (k + 1) for k in range(sockets)]
    if has_gpu:
        n_gpu = int(rng.choice([2, 4, 4, 8])) if flavor != 'wild' else int(rng.integers(1, 9))
        for d in range(n_gpu):
            numa = d * sockets // n_gpu if flavor != 'wild' or rng.random() < 0.8 else int(rng.integers(0, sockets))
            sw = sw_base[numa] + (d % 2 if n_gpu > 2 * sockets else 0)
            if flavor == 'wild' and rng.random() < 0.15:      # NUMA-inconsistent switch
                sw = sw_base[int(rng.integers(0, sockets))]
            gpus.append((d, numa, sw))
    nic_i = 0
    sriov = []
    for k in range(sockets):
        n_k = int(rng.integers(1, 3)) if flavor != 'vf' else 4
        if flavor == 'wild' and rng.random() < 0.1:
            n_k = 0
        for j in range(n_k):
            mbps = int(rng.choice([100000, 100000, 25000, 40000, 10000, 0]) if flavor == 'wild'
                       else rng.choice([100000, 100000, 25000]))
            sw = sw_base[k] + (j % 2 if has_gpu and len(gpus) > 2 * sockets else 0)
            nics.append((f'eth{nic_i}', mbps, k, sw))
            nic_i += 1
    if flavor == 'wild':
        order = rng.permutation(len(nics))
        nics = [nics[i] for i in order]
        if rng.random() < 0.2 and nics:
            sriov = [nics[0][0]]
    groups = None
    if rng.random() < 0.5:
        k = int(rng.integers(1, 3))
        groups = [GROUP_NAMES[i] for i in rng.choice(n_group_names, size=k, replace=False)]
    hp = int(rng.choice([16, 32, 64]))
    return make_node(name, sockets, phys, smt, reserved, gpus, nics, sriov, groups,
                     hp_alloc=hp, hp_free=hp - int(rng.integers(0, 8)),
                     active=bool(rng.random() < 0.95), maintenance=bool(rng.random() < 0.05),
                     isol=bool(rng.random() < 0.9))


def random_pod(rng, flavor='mixed', n_group_names=4, max_groups=3):
    G = int(rng.integers(1, max_groups + 1))
    gpu_pod = rng.random() < 0.5
    pci = gpu_pod and rng.random() < 0.5
    groups = []
    for _ in range(G):
        n_pairs = int(rng.integers(0, 3))
        if flavor == 'wild':
            speeds = [0, 1, 5, 10, 20, 22.5, 45, 50, 90]
        else:
            speeds = [0, 10, 20, 40]
        pairs = [(float(rng.choice(speeds)) if rng.random() < 0.3 else int(rng.choice([0, 10, 20, 40])),
                  int(rng.choice([0, 10, 20]))) for _ in range(n_pairs)]
        n_g = 0
        if gpu_pod:
            n_g = 1 if (pci and flavor != 'wild') else int(rng.integers(0, 3))
        gpus = [int(rng.integers(0, 3)) for _ in range(n_g)]
        groups.append(make_group(pairs, int(rng.integers(0, 5)), gpus, int(rng.integers(0, 3)),
                                 bool(rng.random() < 0.5), bool(rng.random() < 0.5)))
    pg = ['default'] if rng.random() < 0.6 else [GROUP_NAMES[i] for i in
                                                 rng.choice(n_group_names, size=int(rng.integers(1, 3)), replace=False)]
    return make_pod(groups, int(rng.integers(0, 4)), bool(rng.random() < 0.5),
                    'PCI' if pci else 'NUMA', int(rng.choice([0, 0, 2, 4])), pg)


def random_scenario(seed, n_nodes=8, n_pods=24, flavor='mixed', min_busy_secs=30.0, max_groups=3):
    rng = np.random.default_rng(seed)
    nodes = [random_node(rng, f'n{i}', flavor) for i in range(n_nodes)]
    pods = [random_pod(rng, flavor, max_groups=max_groups) for _ in range(n_pods)]
    t = 1000.0
    now = []
    for _ in range(n_pods):
        r = rng.random()
        t += 0.0 if r < 0.5 else (float(rng.integers(1, 20)) if r < 0.9 else 31.0)
        now.append(t)
    return {'nodes': nodes, 'pods': pods, 'now': now, 'min_busy_secs': min_busy_secs}


# ----------------------------------------------------------------------------------
# objects near the packed layout's limits (include/nhd_b200.h: 256 logical cores, 16 GPUs, 32 NICs per
# node; 72 cores per pod; numa^(groups+1) <= 256 tuples)
# ----------------------------------------------------------------------------------
def huge_node(rng, name):
    sockets = int(rng.choice([1, 2, 4]))
    phys = int(rng.choice([64, 96, 128]))
    smt = bool(rng.random() < 0.7)
    n_gpu = int(rng.choice([0, 8, 16]))
    gpus = [(d, d * sockets // max(1, n_gpu), 0x10 * (d * sockets // max(1, n_gpu) + 1) + (d % 4)) for d in range(n_gpu)]
    n_nic = int(rng.choice([2, 8, 16, 32]))
    nics = [(f'vf{i}', int(rng.choice([25000, 100000, 40000])), i % sockets, 0x10 * ((i % sockets) + 1) + (i % 4))
            for i in range(n_nic)]
    return make_node(name, sockets, phys, smt, int(rng.integers(0, 3)), gpus, nics, (), None, hp_alloc=128)


def huge_pod(rng):
    G = int(rng.integers(1, 4))                      # 4 groups on a 4-NUMA node would exceed the tuple limit
    gpu = rng.random() < 0.5
    pci = gpu and rng.random() < 0.5
    groups, budget = [], 72 - 3
    for g in range(G):
        pairs = [(int(rng.choice([0, 5, 10, 20])), int(rng.choice([0, 5, 10]))) for _ in range(int(rng.integers(0, 3)))]
        feeders = [int(rng.integers(0, 3)) for _ in range((1 if pci else int(rng.integers(0, 3))) if gpu else 0)]
        w, h = int(rng.integers(0, 8)), int(rng.integers(0, 4))
        need = 2 * len(pairs) + w + h + sum(feeders)
        if need > budget // (G - g):
            w = max(0, w - (need - budget // (G - g)))
            need = 2 * len(pairs) + w + h + sum(feeders)
        budget -= need
        groups.append(make_group(pairs, w, feeders, h, bool(rng.random() < 0.5), bool(rng.random() < 0.5)))
    return make_pod(groups, int(rng.integers(0, 4)), bool(rng.random() < 0.5), 'PCI' if pci else 'NUMA',
                    int(rng.choice([0, 2, 8])))


def huge_scenario(seed, min_busy_secs=30.0):
    rng = np.random.default_rng(seed)
    n_pods = int(rng.integers(20, 120))
    now = [1000.0] * n_pods if seed % 2 else [float(x) for x in 1000.0 + np.cumsum(rng.integers(0, 12, n_pods))]
    return {'nodes': [huge_node(rng, f'n{i}') for i in range(int(rng.integers(3, 20)))],
            'pods': [huge_pod(rng) for _ in range(n_pods)], 'now': now, 'min_busy_secs': min_busy_secs}
