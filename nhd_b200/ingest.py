"""Node ingest through the library: NFD label dictionaries -> packed ``nhd_node_rec`` records.

The native counterpart of ``Node.ParseLabels`` + ``Node.SetHugepages`` (``nhd/Node.py:468-493``) followed
by ``packing.pack_node``: a cluster can be uploaded (``Solver.load_nodes`` / ``update_nodes``) straight from
the labels the K8s API returns, without building ``Node`` objects first.  The parsing lives in
``csrc/nhd_ingest.cpp`` (C-ABI ``nhd_ingest_*``, ``include/nhd_b200.h``); this module only marshals.
"""
import ctypes
from typing import Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np

from nhd_b200 import _lib, wire


class LabelError(ValueError):
    """``code`` is the library status: ``wire.ERR_LABELS`` where the reference's ``ParseLabels`` returns
    False (the scheduler ignores the node), ``wire.ERR_INVALID`` where it would raise, ``wire.ERR_UNSUPPORTED``
    beyond the packed layout's limits."""

    def __init__(self, code: int, name: str = ''):
        super().__init__(f'node {name!r}: label ingest failed with status {code}')
        self.code = code


class LabelIngest:
    """Holds the cluster-wide dictionaries the packed records refer to (node-group names -> ``group_mask``
    bits, NIC speeds -> speed classes), like ``packing.ClusterLayout`` does for objects."""

    def __init__(self):
        self._L = _lib.load()
        h = ctypes.c_void_p()
        rc = self._L.nhd_ingest_create(ctypes.byref(h))
        if rc != wire.OK:
            raise RuntimeError(f'nhd_ingest_create failed: {rc}')
        self._h = h

    def close(self):
        if self._h:
            self._L.nhd_ingest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def node(self, labels: Mapping[str, str], active: bool = True, hugepages_alloc_gb: int = 0,
             hugepages_free_gb: int = 0, name: str = '', out=None) -> Tuple[np.ndarray, np.ndarray]:
        """One node: ``(nhd_node_rec, nhd_node_aux)``.  ``labels`` in dictionary order; ``out``: a one-element
        slice of a record array to fill in place."""
        n = len(labels)
        keys = (ctypes.c_char_p * n)(*[k.encode() for k in labels.keys()])
        vals = (ctypes.c_char_p * n)(*[str(v).encode() for v in labels.values()])
        rec = np.zeros((), dtype=wire.NODE_DTYPE) if out is None else out
        aux = np.zeros((), dtype=wire.NODE_AUX_DTYPE)
        rc = self._L.nhd_ingest_node(self._h, n, keys, vals, int(bool(active)), int(hugepages_alloc_gb),
                                     int(hugepages_free_gb), rec.ctypes.data, aux.ctypes.data)
        if rc != wire.OK:
            raise LabelError(rc, name)
        return rec, aux

    def nodes(self, items: Sequence[Tuple[Mapping[str, str], bool, int, int]], skip_rejected: bool = False,
              strict: bool = False):
        """Many nodes, as ``NHDScheduler.BuildInitialNodeList`` brings a cluster up (``NHDScheduler.py:83-100``):
        ``items`` = ``(labels, active, hugepages_alloc_gb, hugepages_free_gb)``.  Every node keeps its position in
        the order (the node index is the scheduler's preference order); a node is left in place as an INACTIVE
        record when ``ParseLabels`` would return False (``:87-90``), when it would raise (``except Exception``,
        ``:96-98``), when its allocatable hugepages are 0 (``:92-94``), or when it is beyond the packed limits.
        Returns ``(records, kept)`` with ``kept`` the positions that came up active-capable.

        ``skip_rejected=True`` instead DROPS the nodes ``ParseLabels`` refuses (positions shift; for callers that
        build their own node list); ``strict=True`` re-raises instead of deactivating."""
        recs = np.zeros(len(items), dtype=wire.NODE_DTYPE)
        kept = []
        n_out = 0
        for i, (labels, active, alloc, free) in enumerate(items):
            try:
                self.node(labels, active, alloc, free, name=str(i), out=recs[n_out:n_out + 1])
                if int(alloc) == 0:                                   # NHDScheduler.py:92-94
                    recs[n_out]['flags'] &= ~np.uint8(wire.NODE_ACTIVE)
                else:
                    kept.append(i)
            except LabelError as e:
                if skip_rejected and e.code == wire.ERR_LABELS:
                    continue
                if strict or (skip_rejected and e.code != wire.ERR_LABELS):
                    raise
                recs[n_out] = np.zeros((), dtype=wire.NODE_DTYPE)      # inactive stub, holds the position
                recs[n_out]['n_numa'] = 1
                recs[n_out]['phys_cores'] = 1
            n_out += 1
        return recs[:n_out], kept

    def group_mask(self, names: Iterable[str], create: bool = False) -> int:
        m = ctypes.c_uint64(0)
        rc = self._L.nhd_ingest_group_mask(self._h, '.'.join(names).encode(), int(create), ctypes.byref(m))
        if rc != wire.OK:
            raise LabelError(rc)
        return int(m.value)

    def speed_table(self) -> np.ndarray:
        t = np.zeros(wire.MAX_SPEED_CLASSES, dtype='<f8')
        n = ctypes.c_int32(0)
        rc = self._L.nhd_ingest_speed_table(self._h, t.ctypes.data, ctypes.byref(n))
        if rc != wire.OK:
            raise RuntimeError(f'nhd_ingest_speed_table failed: {rc}')
        return t


def node_stats(recs: np.ndarray) -> np.ndarray:
    """``NHDScheduler.GetBasicNodeStats`` counters (``nhd/NHDScheduler.py:355-378``) of packed records, e.g. of
    ``Solver.read_nodes()``: one ``wire.NODE_STATS_DTYPE`` row per record."""
    recs = np.ascontiguousarray(recs, dtype=wire.NODE_DTYPE).reshape(-1)
    out = np.zeros(len(recs), dtype=wire.NODE_STATS_DTYPE)
    rc = _lib.load().nhd_node_stats_from_records(len(recs), recs.ctypes.data, out.ctypes.data)
    if rc != wire.OK:
        raise ValueError(f'nhd_node_stats_from_records failed: {rc}')
    return out
