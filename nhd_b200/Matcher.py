"""``Matcher`` — drop-in replacement of the reference's ``nhd/Matcher.py`` backed by the
B200 solver.

Same surface as the reference (``Matcher()`` no-arg constructor, ``FindNode(nl, top)``
returning ``(nodename, mapping)`` or ``(None,)``, ``Matcher.py:22-63``), so
``NHDScheduler.AttemptScheduling`` (``NHDScheduler.py:277-304``) can call it unchanged;
plus ``FindNodesBatch`` which schedules a whole pending set (``CheckPendingPods``,
``NHDScheduler.py:425-441``) in one GPU call with the same sequential semantics.

``nl`` may hold this package's ``Node`` mirrors or the reference's own ``nhd.Node.Node``
objects (duck-typed packing).  The filter, the NUMA/PCI intersection, node selection,
mapping choice and the physical core/GPU/NIC assignment all run in CUDA; this module only
converts formats.  Without the CUDA library / a GPU it raises — there is no CPU path.
"""
import time
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from nhd_b200 import packing, wire
from nhd_b200.solver import Solver


class Matcher:
    def __init__(self, device: int = 0):
        self.device = device
        self._solver: Optional[Solver] = None
        self._layout = packing.ClusterLayout()
        self._speeds: List[float] = []
        self._min_busy = None
        self.unsupported_nodes: Dict[str, str] = {}      # active nodes kept out of placement: name -> reason

    # ------------------------------------------------------------------ internals
    def _ensure_solver(self, nodes: Sequence):
        """(Re)create the device handle when the solver constants changed."""
        min_busy = float(getattr(type(nodes[0]), 'MIN_BUSY_SECS', 30.0)) if nodes else 30.0
        if self._solver is None or self._speeds != self._layout.speeds or self._min_busy != min_busy:
            if self._solver is not None:
                self._solver.close()
            self._speeds = list(self._layout.speeds)
            self._min_busy = min_busy
            self._solver = Solver(self._layout.speed_table(), nic_bw_avail_percent=0.9,
                                  min_busy_secs=min_busy, device=self.device)
        return self._solver

    def _load(self, nl: Dict[str, object]):
        names = list(nl.keys())
        nodes = [nl[n] for n in names]
        # may add speed classes / group names; a node outside the packed limits becomes an inactive stub (logged,
        # listed in unsupported_nodes) and the others stay schedulable, as with the reference
        recs = packing.pack_nodes(nodes, self._layout, unsupported=self.unsupported_nodes)
        solver = self._ensure_solver(nodes)
        solver.load_nodes(recs)
        return names, nodes, solver

    # ------------------------------------------------------------------ reference API
    def FindNode(self, nl, top, pod_groups: Iterable[str] = None, now: float = None):
        """Reference signature ``FindNode(nl, top)`` (``Matcher.py:27``).  ``nl`` is the node
        dict already narrowed by ``InitialNodeFilter``; therefore every node's groups match
        by construction unless ``pod_groups`` is given.

        Returns ``(nodename, mapping)`` or ``(None,)``.  ``mapping`` carries the complete
        device-computed binding under ``'_nhd_binding'`` for ``Node.SetPhysicalIdsFromMapping``
        of this package's ``Node`` mirror; the reference's own ``Node`` ignores that key and
        recomputes the identical assignment itself."""
        if len(nl) == 0:
            return (None,)
        names, nodes, solver = self._load(nl)
        pod = packing.pack_pod(top, [], self._layout)
        if pod_groups is None:
            pod['group_mask'] = np.uint64(0xFFFFFFFFFFFFFFFF)   # nl is pre-filtered (NHDScheduler.py:274)
        else:
            pod['group_mask'] = self._layout.groups_mask(pod_groups, create=False)
        now = time.monotonic() if now is None else now
        b = solver.solve_batch(pod.reshape(1), [now])[0]
        if int(b['node']) < 0:
            return (None,)
        mapping = packing.binding_to_mapping(b)
        mapping['_nhd_binding'] = b
        return names[int(b['node'])], mapping

    def FindNodesBatch(self, nl, tops: Sequence, pod_groups: Sequence[Iterable[str]] = None,
                       now=None, apply: bool = True):
        """Schedule ``tops`` in order on ``nl`` (all nodes of the cluster, NOT pre-filtered:
        the per-pod ``InitialNodeFilter`` runs on the GPU from ``pod_groups``).

        Returns one ``(nodename, mapping)`` / ``(None,)`` per pod.  With ``apply`` the bindings
        are also written into the ``Node`` / ``CfgTopology`` objects exactly as the scheduler's
        ``SetBusy`` + ``SetPhysicalIdsFromMapping`` + ``ClaimPodNICResources`` sequence would
        (``NHDScheduler.py:289-304``)."""
        n = len(tops)
        if n == 0:
            return []
        if len(nl) == 0:
            return [(None,)] * n
        names, nodes, solver = self._load(nl)
        pods = np.zeros(n, dtype=wire.POD_DTYPE)
        for i, top in enumerate(tops):
            groups = ['default'] if pod_groups is None else pod_groups[i]
            packing.pack_pod(top, groups, self._layout, out=pods[i])
        if now is None:
            now = time.monotonic()
        nows = np.full(n, now, dtype='<f8') if np.isscalar(now) else np.asarray(now, dtype='<f8')
        bindings = solver.solve_batch(pods, nows)
        results = []
        failed = set()
        for i, (top, b) in enumerate(zip(tops, bindings)):
            if int(b['node']) < 0:
                results.append((None,))
                continue
            node = nodes[int(b['node'])]
            mapping = packing.binding_to_mapping(b)
            mapping['_nhd_binding'] = b
            results.append((names[int(b['node'])], mapping))
            if apply:
                node.busy_time = float(nows[i])                                  # SetBusy
                if int(b['status']) == wire.PLACED:
                    packing.apply_binding(node, top, b)                          # SetPhysicalIdsFromMapping
                    node.ClaimPodNICResources([int(x) for x in b['claimed_nics'][:int(b['n_claimed'])]])
                else:
                    failed.add(int(b['node']))
        # a failed assignment gives cores and GPUs back but not the hugepages taken before the last step
        # (Node.py:794-796 vs :825-837); the solver's records hold the nodes as the reference leaves them
        for idx in failed:
            nodes[idx].mem.free_hugepages_gb = int(solver.read_nodes(idx, 1)[0]['free_hugepages_gb'])
        return results
