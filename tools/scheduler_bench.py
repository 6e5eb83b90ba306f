"""Scheduler-level throughput: `nhd_b200.NHDScheduler.CheckPendingPods` with the whole pending set
as one batch, over an in-memory Kubernetes stand-in (tests/fake_k8s.py) — the per-pod Python work AROUND the solve made
visible: K8s reads + request codec (`_prepare`), record packing + upload of changed nodes (`cluster.sync`), pod packing +
kernels + binding download (`cluster.solve`, device time from the library), `apply_binding` + K8s writes (`_finish`).
Reference counterpart: NHDScheduler.py:249-353, 425-441 (one AttemptScheduling per pod).

    python tools/scheduler_bench.py [n_nodes] [n_pods] [--oracle]     (--oracle: CPU stand-in solver, small sizes only)
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workload
from tests import fake_k8s, helpers, pyref, sched_harness

args = [a for a in sys.argv[1:] if not a.startswith('--')]
n_nodes = int(args[0]) if len(args) > 0 else 65536
n_pods = int(args[1]) if len(args) > 1 else 4096
use_oracle = '--oracle' in sys.argv

import nhd_b200.CfgTopology as cfg_mod
import nhd_b200.Node as node_mod
from nhd_b200.NHDScheduler import NHDScheduler

recs, speed, pods, now = workload.make_workload(4, n_nodes=n_nodes, n_pods=n_pods)
t0 = time.perf_counter()
node_defs = [pyref.node_def_from_record(i, r, speed) for i, r in enumerate(recs)]
k8s = fake_k8s.FakeK8s(node_defs)
setup_s = time.perf_counter() - t0

clock = sched_harness.Clock(1000.0)
sched = NHDScheduler(k8s, lambda cfgtype, cfgstr: fake_k8s.JsonCfgParser(cfgstr, cfg_mod),
                     solver_factory=helpers.OracleSolver if use_oracle else None, clock=clock)
t0 = time.perf_counter()
sched.Startup()                                   # BuildInitialNodeList: ParseLabels per node
startup_s = time.perf_counter() - t0
for name, r in zip(sched.nodes, recs):            # the occupancy of the benchmark cluster, behind the scheduler's back
    pyref.apply_occupancy(sched.nodes[name], r)
    sched._touch(name)

for i, p in enumerate(pods):                      # the pending set appears
    k8s.add_pod('ns', f'pod{i:05d}', pyref.pod_dict_from_record(p))
acc = {}


def timed(obj, attr, key):
    orig = getattr(obj, attr)

    def wrap(*a, **kw):
        t = time.perf_counter()
        try:
            return orig(*a, **kw)
        finally:
            acc[key] = acc.get(key, 0.0) + time.perf_counter() - t
    setattr(obj, attr, wrap)


timed(sched, '_prepare', 'prepare (K8s reads + codec)')
timed(sched, '_finish', 'finish (apply_binding + K8s writes)')
timed(sched.cluster, 'sync', 'sync (pack changed nodes + upload)')
timed(sched.cluster, 'solve', 'solve (pack pods + kernels + download)')
t0 = time.perf_counter()
sched.CheckPendingPods()                          # pods not in pod_state count as pending (NHDScheduler.py:431)
total_s = time.perf_counter() - t0
bound = sum(1 for p in k8s.pods.values() if p['node'])
other = total_s - sum(acc.values())
dev_ms = None
if not use_oracle and sched.cluster._solver is not None:
    t = sched.cluster._solver.timing()
    dev_ms = t['total_ms']
first = {'pods_per_s': n_pods / total_s, 'seconds': {k: round(v, 4) for k, v in acc.items()}, 'seconds_total': round(total_s, 4),
         'note': 'first batch after start-up: every node record is packed from its Python object and uploaded once'}
# steady state: the cluster mirror is resident, the solver's own placements need no upload; a second pending set arrives
acc.clear()
pods2, _ = workload.make_pods(4, n_pods, seed=workload.SEED0 + 77)
for i, p in enumerate(pods2):
    k8s.add_pod('ns', f'late{i:05d}', pyref.pod_dict_from_record(p))
t0 = time.perf_counter()
sched.CheckPendingPods()
total_s = time.perf_counter() - t0
bound = sum(1 for p in k8s.pods.values() if p['node'])
other = total_s - sum(acc.values())
if not use_oracle and sched.cluster._solver is not None:
    dev_ms = sched.cluster._solver.timing()['total_ms']
line = {'metric': 'scheduler-level pods/s (NHDScheduler.CheckPendingPods, one batch, in-memory K8s stand-in)',
        'first_batch': first,
        'value': n_pods / total_s, 'unit': 'pods/s', 'n_nodes': n_nodes, 'n_pods': n_pods, 'bound_total': bound,
        'what': 'second pending set (steady state: device-resident cluster, only host-side changes are re-sent)',
        'seconds': {k: round(v, 4) for k, v in acc.items()}, 'seconds_other': round(other, 4), 'seconds_total': round(total_s, 4),
        'device_ms_last_batch': dev_ms, 'batches': sched.cluster.batches, 'full_loads': sched.cluster.full_loads,
        'startup_s (ParseLabels per node, once)': round(startup_s, 2), 'setup_s (bench only)': round(setup_s, 2),
        'solver': 'oracle stand-in (CPU)' if use_oracle else 'nhd_b200 CUDA library'}
print(json.dumps(line))
sched.close()
