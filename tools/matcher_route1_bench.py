"""Route 1 of INTEGRATION.md, per pod: ``Matcher.FindNode(nl, top)`` called the way the unmodified reference scheduler
calls it (``NHDScheduler.py:235-247,277-304``: filter the node dict, FindNode, SetBusy, SetPhysicalIdsFromMapping,
ClaimPodNICResources) on this package's ``Node`` mirrors, with and without the change counters of
``nhd_b200/tracking.py``.  Prints the host time per call split into bringing the device cluster up to date
(pack + upload of what changed) and the rest (pod packing, the solve, the mapping dict).

    python tools/matcher_route1_bench.py [n_nodes] [n_pods] [--null | --oracle]

``--null``: a solver stand-in that answers "no candidate" at once — the Python overhead alone, on a machine without a
GPU; ``--oracle``: the C oracle as the solver (small clusters).  Default: the CUDA library.  Developer tool.
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workload
from tests import helpers, pyref, scenarios
import nhd_b200.CfgTopology as cfg_mod
import nhd_b200.Node as node_mod
import nhd_b200.Matcher as matcher_mod
from nhd_b200 import wire

args = [a for a in sys.argv[1:] if not a.startswith('--')]
n_nodes = int(args[0]) if len(args) > 0 else 65536
n_pods = int(args[1]) if len(args) > 1 else 64


class NullSolver(helpers.OracleSolver):
    def solve_batch(self, pods, now):
        out = np.zeros(len(pods), dtype=wire.BINDING_DTYPE)
        out['node'] = -1
        out['status'] = wire.NO_CANDIDATE
        return out


if '--null' in sys.argv:
    matcher_mod.Solver = NullSolver
elif '--oracle' in sys.argv:
    matcher_mod.Solver = helpers.OracleSolver

recs, speed, pods, now = workload.make_workload(4, n_nodes=n_nodes, n_pods=n_pods)
t0 = time.perf_counter()
nodes = pyref.nodes_from_records(recs, speed, node_mod)
build_s = time.perf_counter() - t0

out = {'n_nodes': n_nodes, 'n_pods': n_pods, 'build_node_objects_s': round(build_s, 2),
       'solver': 'null stand-in' if '--null' in sys.argv else 'oracle stand-in' if '--oracle' in sys.argv else 'CUDA library'}
for track in (True, False):
    m = matcher_mod.Matcher(track=track)
    sync_s, total_s, found = [], [], 0
    orig = m._load

    def timed(nl, _o=orig):
        t = time.perf_counter()
        try:
            return _o(nl)
        finally:
            sync_s.append(time.perf_counter() - t)
    m._load = timed
    t_now = 1000.0
    for i in range(n_pods if track else min(n_pods, 6)):
        top = scenarios.build_top(pyref.pod_dict_from_record(pods[i]), cfg_mod)
        nl = {k: v for k, v in nodes.items() if v.active and v.groups}          # InitialNodeFilter
        t = time.perf_counter()
        match = m.FindNode(nl, top, now=t_now)
        total_s.append(time.perf_counter() - t)
        t_now += 1.0
        if match != (None,):
            found += 1
            node = nodes[match[0]]
            node.busy_time = t_now
            try:
                nic_list = node.SetPhysicalIdsFromMapping(match[1], top)
                node.ClaimPodNICResources(list({x[0] for x in nic_list}))
            except IndexError:
                pass
    key = 'change_counters' if track else 'pack_everything'
    out[key] = {'calls': len(total_s), 'placed': found,
                'first_call_ms': round(1e3 * total_s[0], 2),
                'later_calls_ms': round(1e3 * float(np.mean(total_s[1:])), 2),
                'of_which_cluster_sync_ms': round(1e3 * float(np.mean(sync_s[1:])), 2),
                'stats': dict(m.stats)}
    if m._solver is not None:
        m._solver.close()
print(json.dumps(out))
