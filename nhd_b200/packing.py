"""Host glue between the reference-shaped Python objects and the packed wire records.

``pack_node`` / ``pack_pod`` read *duck-typed* objects: either this package's mirrors
(``nhd_b200.Node.Node``, ``nhd_b200.CfgTopology.CfgTopology``) or the reference's own
``nhd.Node.Node`` / ``nhd.CfgTopology.CfgTopology`` instances, so the solver can be
dropped under an unmodified ``NHDScheduler``.  ``apply_binding`` is the inverse: it
writes a solver binding back where ``Node.SetPhysicalIdsFromMapping``
(``nhd/Node.py:663-841``) would have left it.

No placement decisions are taken here — only format conversion.
"""
import logging
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from nhd_b200 import tracking, wire


_log = logging.getLogger('nhd_b200')


class UnsupportedError(ValueError):
    """The object is valid for the reference but outside the packed layout's limits
    (include/nhd_b200.h, NHD_MAX_*)."""


class ClusterLayout:
    """Cluster-wide dictionaries the packed records refer to.

    * node-group names (``Node.groups`` / pod annotation ``nhd_groups``,
      ``NHDScheduler.py:235-247``) -> bit position in ``group_mask``;
    * NIC link speeds (``NodeNic.speed``, Gb/s) -> 4-bit speed class.
    """

    def __init__(self):
        self.group_bits: Dict[str, int] = {}
        self.speeds: List[float] = []

    def group_bit(self, name: str, create: bool = True) -> Optional[int]:
        bit = self.group_bits.get(name)
        if bit is None and create:
            if len(self.group_bits) >= 64:
                raise UnsupportedError('more than 64 distinct node-group names')
            bit = self.group_bits[name] = len(self.group_bits)
        return bit

    def groups_mask(self, names: Iterable[str], create: bool) -> int:
        mask = 0
        for n in names:
            bit = self.group_bit(n, create)
            if bit is not None:
                mask |= 1 << bit
        return mask

    def speed_class(self, speed: float) -> int:
        speed = float(speed)
        for i, s in enumerate(self.speeds):
            if s == speed:
                return i
        if len(self.speeds) >= wire.MAX_SPEED_CLASSES:
            raise UnsupportedError(f'more than {wire.MAX_SPEED_CLASSES} distinct NIC speeds')
        self.speeds.append(speed)
        return len(self.speeds) - 1

    def speed_table(self) -> np.ndarray:
        t = np.zeros(wire.MAX_SPEED_CLASSES, dtype='<f8')
        t[:len(self.speeds)] = self.speeds
        return t


def _enum_value(x) -> int:
    return x if type(x) is int else int(getattr(x, 'value', x))


def pack_node(node, layout: ClusterLayout, out=None):
    """One ``Node`` (reference or mirror) -> one ``nhd_node_rec``."""
    rec = np.zeros((), dtype=wire.NODE_DTYPE) if out is None else out
    K = int(node.numa_nodes)
    if not 1 <= K <= wire.MAX_NUMA:
        raise UnsupportedError(f'node {node.name}: {K} NUMA nodes (max {wire.MAX_NUMA})')
    smt = bool(node.smt_enabled)
    n_logical = len(node.cores)
    phys = n_logical // 2 if smt else n_logical
    if n_logical > wire.MAX_LCORES:
        raise UnsupportedError(f'node {node.name}: {n_logical} logical cores (max {wire.MAX_LCORES})')
    if phys == 0 or phys % K != 0:
        raise UnsupportedError(f'node {node.name}: {phys} physical cores not divisible by {K} sockets')
    per_sock = phys // K
    used = [0, 0, 0, 0]
    for i, c in enumerate(node.cores):
        if c.socket != (i % phys) // per_sock or (smt and c.sibling != (i + phys if i < phys else i - phys)):
            raise UnsupportedError(f'node {node.name}: irregular core {i} topology')
        if c.used:
            used[i >> 6] |= 1 << (i & 63)
    rec['used'] = used
    rec['n_numa'] = K
    rec['phys_cores'] = phys
    rec['flags'] = ((wire.NODE_SMT if smt else 0) | (wire.NODE_ACTIVE if node.active else 0)
                    | (wire.NODE_MAINTENANCE if node.maintenance else 0))
    rec['free_hugepages_gb'] = int(node.mem.free_hugepages_gb)
    rec['busy_time'] = float(node.busy_time)
    rec['group_mask'] = layout.groups_mask(node.groups, create=True)

    switches: Dict[int, int] = {}

    def local_switch(sw):
        if sw not in switches:
            if len(switches) >= wire.MAX_SWITCHES:
                raise UnsupportedError(f'node {node.name}: more than {wire.MAX_SWITCHES} PCIe switches')
            switches[sw] = len(switches)
        return switches[sw]

    if len(node.gpus) > wire.MAX_GPUS:
        raise UnsupportedError(f'node {node.name}: {len(node.gpus)} GPUs (max {wire.MAX_GPUS})')
    gpu_used = 0
    gpu_numa = [0] * wire.MAX_NUMA
    gpu_sw = 0
    for i, g in enumerate(node.gpus):
        if not 0 <= g.numa_node < K:
            raise UnsupportedError(f'node {node.name}: GPU {i} on NUMA node {g.numa_node} of {K}')
        # the reference's failure unwind frees self.gpus[device_id] (Node.py:828): only exact while ids are positions
        if int(getattr(g, 'device_id', i)) != i:
            raise UnsupportedError(f'node {node.name}: GPU {i} carries device id {g.device_id} (ids must equal list positions)')
        gpu_numa[g.numa_node] |= 1 << i
        gpu_sw |= local_switch(g.pciesw) << (4 * i)
        if g.used:
            gpu_used |= 1 << i
    rec['n_gpus'] = len(node.gpus)
    rec['gpu_used'] = gpu_used
    rec['gpu_numa_mask'] = gpu_numa
    rec['gpu_sw'] = gpu_sw

    if len(node.nics) > wire.MAX_NICS:
        raise UnsupportedError(f'node {node.name}: {len(node.nics)} NICs (max {wire.MAX_NICS})')
    nic_numa = [0] * wire.MAX_NUMA
    nic_sw = [0, 0]
    nic_speed = [0, 0]
    nic_inuse = 0
    running = [0] * wire.MAX_NUMA
    for i, n in enumerate(node.nics):
        if not 0 <= n.numa_node < K:
            raise UnsupportedError(f'node {node.name}: NIC {i} on NUMA node {n.numa_node} of {K}')
        if n.idx != running[n.numa_node]:
            raise UnsupportedError(f'node {node.name}: NIC {i} has idx {n.idx}, expected {running[n.numa_node]}')
        running[n.numa_node] += 1
        nic_numa[n.numa_node] |= 1 << i
        nic_sw[i >> 4] |= local_switch(n.pciesw) << (4 * (i & 15))
        nic_speed[i >> 4] |= layout.speed_class(n.speed) << (4 * (i & 15))
        if n.pods_used > 0:
            nic_inuse |= 1 << i
    rec['n_nics'] = len(node.nics)
    rec['nic_numa_mask'] = nic_numa
    rec['nic_sw'] = nic_sw
    rec['nic_speed'] = nic_speed
    rec['nic_inuse'] = nic_inuse
    return rec


def stub_record(out):
    """Inactive place-holder record: keeps a node's position in the order, is never a candidate."""
    for name in wire.NODE_DTYPE.names:                  # `out` may be a record of an array (numpy.void): field by field
        out[name] = 0
    out['n_numa'] = 1
    out['phys_cores'] = 1
    return out


def pack_nodes(nodes: Sequence, layout: ClusterLayout, unsupported: dict = None) -> np.ndarray:
    """Records of ``nodes`` in order.  A node the packed layout cannot describe (beyond ``NHD_MAX_*``, irregular
    core / NIC / GPU numbering) becomes an inactive stub at its position: the remaining nodes stay schedulable, as
    they would with the reference, instead of the whole call failing.  Such nodes are logged, and reported through
    ``unsupported`` (name -> reason) when given."""
    recs = np.zeros(len(nodes), dtype=wire.NODE_DTYPE)
    for i, n in enumerate(nodes):
        try:
            pack_node(n, layout, out=recs[i])
        except UnsupportedError as err:
            stub_record(recs[i])
            if getattr(n, 'active', True):
                _log.warning('node %s is outside the packed limits and is kept out of placement: %s',
                             getattr(n, 'name', i), err)
                if unsupported is not None:
                    unsupported[getattr(n, 'name', str(i))] = str(err)
    return recs


def pack_pod(top, pod_groups: Iterable[str], layout: ClusterLayout, out=None):
    """One ``CfgTopology`` (+ the pod's node-group list) -> one ``nhd_pod``.

    Follows the request vectors of ``CfgTopology.py:199-232``: per group the GPU count,
    proc cores + GPU feeder cores, helper cores, SMT flags and RX/TX Gb/s sums."""
    rec = np.zeros((), dtype=wire.POD_DTYPE) if out is None else out
    G = len(top.proc_groups)
    if not 1 <= G <= wire.MAX_GROUPS:
        raise UnsupportedError(f'{G} processing groups (supported 1..{wire.MAX_GROUPS})')
    rec['n_groups'] = G
    rec['map_type'] = _enum_value(top.map_type)
    n_misc = len(top.misc_cores)
    rec['n_misc'] = n_misc
    rec['flags'] = wire.POD_MISC_SMT if _enum_value(top.misc_cores_smt) else 0
    rec['hugepages_gb'] = int(top.hugepages_gb)
    rec['group_mask'] = layout.groups_mask(pod_groups, create=False)
    total_cores, total_gpus = n_misc, 0
    for gi, pg in enumerate(top.proc_groups):
        g = rec['groups'][gi]
        ngpu = len(pg.group_gpus)
        if ngpu > wire.MAX_GROUP_GPUS:
            raise UnsupportedError(f'group {gi}: {ngpu} GPUs (max {wire.MAX_GROUP_GPUS})')
        feeders = [len(gpu.cpu_cores) for gpu in pg.group_gpus]
        g['n_gpus'] = ngpu
        g['gpu_feeders'][:ngpu] = feeders
        g['n_proc'] = len(pg.proc_cores)
        g['n_helpers'] = len(pg.misc_cores)
        rx = tx = 0
        has_nic = False
        for c in pg.proc_cores:
            d = _enum_value(c.nic_dir)
            if d == 1:
                rx += c.nic_speed
                has_nic = True
            elif d == 2:
                tx += c.nic_speed
                has_nic = True
        g['rx_gbps'] = float(rx)
        g['tx_gbps'] = float(tx)
        g['flags'] = ((wire.GRP_PROC_SMT if _enum_value(pg.proc_smt) else 0)
                      | (wire.GRP_HELPER_SMT if _enum_value(pg.helper_smt) else 0)
                      | (wire.GRP_HAS_NIC_CORES if has_nic else 0))
        total_cores += len(pg.proc_cores) + len(pg.misc_cores) + sum(feeders)
        total_gpus += ngpu
    if total_cores > wire.MAX_POD_CORES:
        raise UnsupportedError(f'pod requests {total_cores} cores (max {wire.MAX_POD_CORES})')
    if total_gpus > wire.MAX_POD_GPUS:
        raise UnsupportedError(f'pod requests {total_gpus} GPUs (max {wire.MAX_POD_GPUS})')
    return rec


def binding_to_mapping(b) -> dict:
    """``nhd_binding`` -> the dict ``Matcher.FindNode`` returns (``Matcher.py:452``)."""
    G = int(b['n_groups'])
    return {'gpu': tuple(int(x) for x in b['gpu_numa'][:G]),
            'cpu': tuple(int(x) for x in b['cpu_numa'][:G + 1]),
            'nic': [(int(b['nic_numa'][g]), int(b['nic_idx'][g])) for g in range(G)]}


def apply_binding(node, top, b):
    """Write a solver binding into ``top`` / ``node`` the way
    ``Node.SetPhysicalIdsFromMapping`` does (``Node.py:674-818``) and return ``used_nics``.
    Raises ``IndexError`` for a failed assignment (``Node.py:825-837``)."""
    status = int(b['status'])
    tracking.bump(node)                                  # the node's sub-objects change below (Matcher's change counter)
    if status in (wire.ASSIGN_FAILED, wire.REF_WOULD_CRASH):
        raise IndexError('physical assignment failed on node %s' % node.name)
    if status != wire.PLACED:
        raise RuntimeError('binding has no placement (status %d)' % status)
    cores = b['cores'][:int(b['n_cores'])].tolist()
    gpus = b['gpu_index'][:int(b['n_gpus'])].tolist()
    nic_list_index = b['nic_list_index'].tolist()
    ci = gi = 0
    used_nics = []
    for pi, pv in enumerate(top.proc_groups):
        if pv.vlan is not None:
            pv.vlan.vlan = node.data_vlan
        nic_index = nic_list_index[pi]
        nic = node.nics[nic_index]
        for gv in pv.group_gpus:
            dev = node.gpus[gpus[gi]]
            gi += 1
            gv.device_id = dev.device_id
            dev.used = True
            for c in gv.cpu_cores:
                c.core = cores[ci]
                node.cores[c.core].used = True
                ci += 1
        for c in pv.proc_cores:
            c.core = cores[ci]
            node.cores[c.core].used = True
            ci += 1
            d = getattr(c.nic_dir, 'value', c.nic_dir)
            if d in (1, 2):
                nic.speed_used[0 if d == 1 else 1] += c.nic_speed
                used_nics.append((nic_index, c.nic_speed, c.nic_dir))
                pair = top.GetNICGroup(c)
                if pair is None:
                    raise IndexError('core %s is not part of a NIC pairing' % c.name)
                pair.AddInterface(nic.mac)
        for c in pv.misc_cores:
            c.core = cores[ci]
            node.cores[c.core].used = True
            ci += 1
    top.SetDataDefaultGw(node.gwip)
    if top.hugepages_gb > 0:
        node.mem.free_hugepages_gb -= top.hugepages_gb
    for c in top.misc_cores:
        c.core = cores[ci]
        node.cores[c.core].used = True
        ci += 1
    if top.ctrl_vlan is not None:
        top.ctrl_vlan.vlan = node.data_vlan
    return used_nics
