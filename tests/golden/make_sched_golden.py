"""Freeze scheduler sessions played by the UNMODIFIED reference ``nhd.NHDScheduler.run()``
(build container only) into tests/golden/sched/*.json: the script (cluster, Kubernetes
operations, watch items, RPCs, restarts — ``tests/sched_harness.py``) and the state document the
reference ends in (what Kubernetes saw, ``pod_state``, every node's resources, RPC answers).
The GPU box replays the scripts through ``nhd_b200.NHDScheduler`` + the CUDA solver
(``tests/test_scheduler.py``).

    python tests/golden/make_sched_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import sched_harness as H          # noqa: E402
from tests import scenarios                   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'sched')


def unwind_script():
    """Hand-written: one pending set whose Kubernetes writes fail at every step the reference
    unwinds from (NHDScheduler.py:308-343), between pods that succeed."""
    nodes = [scenarios.make_node(f'n{i}', 2, 16, True, 1,
                                 gpus=[(d, d // 2, 0x10 * (d // 2 + 1)) for d in range(4)] if i % 3 != 2 else (),
                                 nics=[('eth0', 100000, 0, 0x10), ('eth1', 100000, 1, 0x20)]) for i in range(12)]

    def pod(gpu=False, hp=2):
        g = scenarios.make_group(pairs=((10, 10),), workers=2, gpus=(1,) if gpu else (), helpers=1)
        return scenarios.make_pod([g], misc=1, hugepages=hp, map_type='PCI' if gpu else 'NUMA')

    init = []
    fails = [[], [], ['bind'], [], ['nad'], [], ['gpumap'], [], ['annotate'], [], ['podobj'], ['cfg'], [], []]
    for i, f in enumerate(fails):
        init.append({'op': 'add_pod', 'ns': 'prod', 'name': f'p{i}', 'uid': f'u{i}', 'fail': f,
                     'pod': pod(gpu=(i % 2 == 0))})
    steps = [{'op': 'rpc', 'msg': 'TYPE_NODE_INFO'}, {'op': 'clock', 'dt': 40.0}, {'op': 'idle'},
             {'op': 'watch', 'type': 'NHD_WATCH_TYPE_TRIAD_POD_DELETE', 'pod': {'ns': 'prod', 'name': 'p1', 'uid': 'u1'}},
             {'op': 'delete_pod', 'ns': 'prod', 'name': 'p1'},
             {'op': 'watch', 'type': 'NHD_WATCH_TYPE_NODE_CORDON', 'node': 'n1'},
             {'op': 'restart'}, {'op': 'clock', 'dt': 40.0}, {'op': 'idle'},
             {'op': 'rpc', 'msg': 'TYPE_NODE_INFO'}, {'op': 'rpc', 'msg': 'TYPE_SCHEDULER_INFO'}]
    return {'nodes': nodes, 'min_busy_secs': 30.0, 'clock0': 1000.0, 'init': init, 'steps': steps}


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = [('unwind_every_step', unwind_script())]
    for seed, flavor in ((1, 'mixed'), (2, 'mixed'), (5, 'mixed'), (3, 'vf'), (7, 'vf'), (4, 'big'), (9, 'big'),
                         (6, 'wild'), (11, 'wild')):
        cases.append((f'random{seed}_{flavor}', H.random_script(seed, flavor)))
    # sessions in which assignments fail after a successful filter (rewind path), and sessions whose pods carry
    # real libconfig text read by the reference's own TriadCfgParser (on the libconf / magicattr stand-ins)
    for seed in (10, 17):
        cases.append((f'random{seed}_wild_assign_failures', H.random_script(seed, 'wild')))
    for seed, flavor in ((2, 'wild'), (3, 'mixed'), (8, 'vf'), (12, 'big')):
        cases.append((f'triad{seed}_{flavor}', H.random_script(seed, flavor, codec='triad')))
    for name, script in cases:
        doc = {'generator': 'tests/golden/make_sched_golden.py', 'source': 'unmodified nhd.NHDScheduler.run()',
               'script': script, 'expected': H.run_reference(script)}
        with open(os.path.join(OUT, name + '.json'), 'w') as f:
            json.dump(doc, f, separators=(',', ':'))      # key order is data: NIC / GPU order follows label order
        print(name, 'binds', len(doc['expected']['k8s']['binds']),
              'failed_schedule_count', doc['expected']['failed_schedule_count'])


if __name__ == '__main__':
    main()
