"""Frozen vectors from the unmodified reference (tests/golden/*.json, made by make_golden.py).

CPU:  the oracle and the host-compiled product core reproduce every recorded outcome.
GPU:  the CUDA solver does, through the C-ABI and through the object-level drop-in API."""
import glob
import json
import os

import numpy as np
import pytest

import nhd_b200.CfgTopology as cfg_mod
import nhd_b200.Node as node_mod
from nhd_b200 import packing, wire
from tests import helpers, ref_compare, scenarios

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.json')))
assert len(GOLDEN) >= 16


def _load(path):
    with open(path) as f:
        return json.load(f)


def _check_bindings(doc, bindings, final_recs):
    for i, (o, b) in enumerate(zip(doc['outcomes'], bindings)):
        errs = ref_compare.diff_outcome(o, b)
        assert not errs, (i, errs)
    # final node state, field by field
    for rec, (name, st) in zip(final_recs, doc['final_state'].items()):
        used = [c for c in range(256) if (int(rec['used'][c >> 6]) >> (c & 63)) & 1]
        assert used == st['used'], name
        assert [i for i in range(16) if (int(rec['gpu_used']) >> i) & 1] == st['gpu_used'], name
        assert [int((int(rec['nic_inuse']) >> i) & 1) for i in range(len(st['pods_used']))] == \
            [int(p > 0) for p in st['pods_used']], name
        assert int(rec['free_hugepages_gb']) == st['free_hugepages'], name
        assert float(rec['busy_time']) == st['busy_time'], name


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_oracle_reproduces_golden(oracle_lib, path):
    doc = _load(path)
    scn = doc['scenario']
    recs, pods, now, layout = ref_compare.pack_scenario(scn)
    b, final = oracle_lib.solve(recs, layout.speed_table(), pods, now, min_busy_secs=scn['min_busy_secs'])
    _check_bindings(doc, b, final)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_product_core_reproduces_golden(emu, path):
    doc = _load(path)
    scn = doc['scenario']
    recs, pods, now, layout = ref_compare.pack_scenario(scn)
    b, final = helpers.emu_solve(emu, recs, layout.speed_table(), pods, now, min_busy=scn['min_busy_secs'])
    _check_bindings(doc, b, final)


@pytest.mark.gpu
@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_cuda_reproduces_golden(path):
    from nhd_b200.solver import Solver
    doc = _load(path)
    scn = doc['scenario']
    recs, pods, now, layout = ref_compare.pack_scenario(scn)
    s = Solver(layout.speed_table(), min_busy_secs=scn['min_busy_secs'])
    try:
        s.load_nodes(recs)
        b = s.solve_batch(pods, now)
        final = s.read_nodes()
    finally:
        s.close()
    _check_bindings(doc, b, final)


@pytest.mark.gpu
@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_dropin_matcher_objects_match_golden(path):
    """Object-level drop-in: nhd_b200.Matcher.FindNodesBatch on Node / CfgTopology mirrors leaves
    every Core.core, GPU.device_id, NICGroup.mac, VLAN and every node counter exactly where the
    reference's scheduler loop leaves them."""
    from nhd_b200.Matcher import Matcher
    doc = _load(path)
    scn = doc['scenario']
    node_mod.Node.MIN_BUSY_SECS = float(scn['min_busy_secs'])
    try:
        nodes = scenarios.build_nodes(scn, node_mod)
        tops = [scenarios.build_top(p, cfg_mod) for p in scn['pods']]
        m = Matcher()
        res = m.FindNodesBatch(nodes, tops, [p['groups'] for p in scn['pods']], now=scn['now'])
    finally:
        node_mod.Node.MIN_BUSY_SECS = 30.0
    names = list(nodes.keys())
    for i, (o, r, top) in enumerate(zip(doc['outcomes'], res, tops)):
        if o['status'] == 'none':
            assert r == (None,), i
            continue
        assert r[0] == names[o['node']], i
        assert {k: [list(x) if isinstance(x, tuple) else x for x in v] if k == 'nic' else list(v)
                for k, v in r[1].items() if k != '_nhd_binding'} == o['mapping'], i
        if o['status'] == 'placed':
            got = scenarios.extract_result(top)
            for k in ('cores', 'gpus', 'macs', 'vlans', 'ctrl_vlan', 'gw'):
                assert got[k] == o[k], (i, k, got[k], o[k])
    for name, st in doc['final_state'].items():
        assert scenarios.node_state(nodes[name]) == st, name


@pytest.mark.gpu
def test_dropin_findnode_single_call_contract():
    """Matcher.FindNode keeps the reference's calling convention (Matcher.py:27-63,
    NHDScheduler.py:277-299): (name, mapping) / (None,), mapping feeds SetPhysicalIdsFromMapping."""
    from nhd_b200.Matcher import Matcher
    doc = _load([p for p in GOLDEN if 'kat1' in p][0])
    scn = doc['scenario']
    node_mod.Node.MIN_BUSY_SECS = 0.0
    try:
        nodes = scenarios.build_nodes(scn, node_mod)
        m = Matcher()
        for o, pod in zip(doc['outcomes'], scn['pods']):
            top = scenarios.build_top(pod, cfg_mod)
            match = m.FindNode(nodes, top, now=1000.0)
            assert match[0] == list(nodes)[o['node']]
            node = nodes[match[0]]
            node.busy_time = 1000.0
            nic_list = node.SetPhysicalIdsFromMapping(match[1], top)
            nidx = list({x[0] for x in nic_list})
            node.ClaimPodNICResources(nidx)
            assert nidx == o['nidx']
            assert scenarios.extract_result(top)['cores'] == o['cores']
        with pytest.raises(RuntimeError):
            node.SetPhysicalIdsFromMapping({'gpu': (0,), 'cpu': (0, 1), 'nic': [(0, 0)]}, top)   # no CPU fallback
        assert m.FindNode({}, top) == (None,)
    finally:
        node_mod.Node.MIN_BUSY_SECS = 30.0
