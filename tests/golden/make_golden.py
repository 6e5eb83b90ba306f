"""Generates tests/golden/*.json from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Each file holds a scenario (node labels, pod specs, clocks) and what the reference's own
Matcher.FindNode + Node.SetPhysicalIdsFromMapping + ClaimPodNICResources sequence
(NHDScheduler.py:277-304) produced for it, pod by pod, plus the final state of every node.
The reference's tests hold no vectors for this path, so these files
are the frozen pins; the known-answer cases follow SURVEY.md appendix D.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import ref_compare, scenarios as S   # noqa: E402

STD_GPUS = [(0, 0, 0x10), (1, 0, 0x10), (2, 1, 0x20), (3, 1, 0x20)]
STD_NICS = [('eth2', 100000, 0, 0x10), ('eth3', 100000, 1, 0x20)]


def std_node(name, gpus=True, nics=STD_NICS, **kw):
    return S.make_node(name, 2, 16, True, gpus=STD_GPUS if gpus else (), nics=nics,
                       isolcpus='2-15_18-31', hp_alloc=64, **kw)


def kat_scenarios():
    out = {}
    pod1 = S.make_pod([S.make_group([(10, 10)], 1, [1], 1)], misc=1)
    out['kat1_gpu_pod_packing'] = dict(nodes=[std_node(f'n{i}') for i in range(4)], pods=[pod1] * 6,
                                       now=[1000.0] * 6, min_busy_secs=0.0)
    out['kat2_busy_window'] = dict(nodes=[std_node(f'n{i}') for i in range(3)], pods=[pod1] * 4,
                                   now=[1000.0] * 4, min_busy_secs=30.0)
    pod3 = S.make_pod([S.make_group([(10, 10)], 2)], misc=1)
    out['kat3_cpu_only_prefers_gpuless'] = dict(
        nodes=[std_node('g0'), std_node('g1'), std_node('c0', gpus=False), std_node('c1', gpus=False)],
        pods=[pod3] * 6, now=[1000.0] * 6, min_busy_secs=0.0)
    nics4 = [('eth2', 100000, 0, 0x10), ('eth3', 100000, 0, 0x10), ('eth4', 100000, 1, 0x20), ('eth5', 100000, 1, 0x20)]
    pod4 = S.make_pod([S.make_group([(10, 10)], 0, [1]), S.make_group([(10, 10)], 0, [1])], misc=1)
    out['kat4_two_groups'] = dict(nodes=[std_node(f'n{i}', nics=nics4) for i in range(2)], pods=[pod4] * 3,
                                  now=[1000.0] * 3, min_busy_secs=0.0)
    pod5 = S.make_pod([S.make_group([(10, 10)], 3, [], 3, proc_smt=True, helper_smt=True)], misc=2, misc_smt=True)
    out['kat5_smt_request'] = dict(nodes=[std_node('n0', gpus=False)], pods=[pod5], now=[1000.0], min_busy_secs=0.0)
    gpus6 = [(i, i // 4, 0x10 * (i // 4 + 1) + (i // 2) % 2) for i in range(8)]
    nics6 = [('eth0', 100000, 0, 0x10), ('eth1', 100000, 0, 0x11), ('eth2', 100000, 1, 0x20), ('eth3', 100000, 1, 0x21)]
    node6 = S.make_node('n0', 2, 32, True, 2, gpus6, nics6, hp_alloc=64)
    pod6 = S.make_pod([S.make_group([(10, 10)], 1, [1]) for _ in range(3)], misc=1, map_type='PCI')
    out['kat6_three_groups_pci'] = dict(nodes=[node6], pods=[pod6] * 3, now=[1000.0] * 3, min_busy_secs=0.0)
    node7 = S.make_node('n0', 2, 8, True, gpus=(), nics=[('eth0', 100000, 0, 0x10), ('eth1', 100000, 1, 0x20)],
                        isolcpus='1-3,9-11,4,12', hp_alloc=64)
    pod7 = S.make_pod([S.make_group([(10, 10)], 1)], misc=2, misc_smt=False)
    out['kat7_misc_hyperthreads'] = dict(nodes=[node7], pods=[pod7], now=[1000.0], min_busy_secs=0.0)
    pod8a = S.make_pod([S.make_group([(91, 1)], 0)], misc=1)
    pod8b = S.make_pod([S.make_group([(45, 45), (45, 45)], 0)], misc=1)
    out['kat8_nic_capacity'] = dict(nodes=[std_node('n0', gpus=False)], pods=[pod8a, pod8b], now=[1000.0] * 2,
                                    min_busy_secs=0.0)
    return out


# expectations quoted in SURVEY.md appendix D (probed on the reference during the survey)
EXPECT = {
    'kat1_gpu_pod_packing': lambda o: (
        o[0]['node'] == 0 and o[0]['mapping'] == {'gpu': [0], 'cpu': [0, 1], 'nic': [[0, 0]]}
        and o[0]['gpus'] == [0] and o[0]['cores'] == [2, 3, 4, 5, 6, 8] and o[0]['nidx'] == [0]
        and o[1]['node'] == 0 and o[1]['mapping'] == {'gpu': [1], 'cpu': [1, 0], 'nic': [[1, 0]]}
        and o[1]['gpus'] == [2] and o[1]['cores'] == [9, 10, 11, 12, 13, 7] and o[1]['nidx'] == [1]
        and [x['node'] for x in o] == [0, 0, 1, 1, 2, 2]),
    'kat2_busy_window': lambda o: [x.get('node') for x in o] == [0, 1, 2, None],
    'kat3_cpu_only_prefers_gpuless': lambda o: (
        [x['node'] for x in o] == [2, 2, 3, 3, 0, 0] and o[0]['cores'] == [2, 3, 4, 5, 8]
        and o[1]['cores'] == [9, 10, 11, 12, 6]),
    'kat4_two_groups': lambda o: (
        o[0]['mapping'] == {'gpu': [1, 1], 'cpu': [1, 1, 0], 'nic': [[1, 0], [1, 0]]} and o[0]['gpus'] == [2, 3]
        and o[0]['cores'] == [8, 9, 10, 11, 12, 13, 2] and o[1]['node'] == 1 and o[2]['status'] == 'none'),
    'kat5_smt_request': lambda o: o[0]['cores'] == [2, 18, 3, 19, 4, 5, 21, 6, 8, 24],
    'kat6_three_groups_pci': lambda o: (
        o[0]['mapping'] == {'gpu': [0, 0, 0], 'cpu': [0, 0, 0, 1], 'nic': [[0, 0], [0, 0], [0, 1]]}
        and o[0]['gpus'] == [0, 1, 2] and o[1]['mapping']['gpu'] == [1, 1, 1] and o[1]['mapping']['cpu'] == [1, 1, 1, 0]
        and o[1]['gpus'] == [4, 5, 6] and o[2]['status'] == 'none'),
    'kat7_misc_hyperthreads': lambda o: o[0]['cores'] == [1, 2, 3, 4, 12],
    'kat8_nic_capacity': lambda o: o[0]['status'] == 'none' and o[1]['status'] == 'placed' and o[1]['nidx'] == [0],
}


def main():
    scns = kat_scenarios()
    for i, (flavor, n_nodes, n_pods, mg) in enumerate([('mixed', 6, 40, 3), ('mixed', 10, 60, 4), ('wild', 6, 40, 3),
                                                       ('wild', 8, 60, 3), ('vf', 5, 40, 3), ('big', 6, 60, 4),
                                                       ('wild', 12, 80, 3), ('mixed', 16, 4, 2)]):
        scns[f'random{i}_{flavor}'] = S.random_scenario(31337 + i, n_nodes, n_pods, flavor, max_groups=mg)
    for name, scn in scns.items():
        outs, init, final, layout, names = ref_compare.run_reference(scn)
        if name in EXPECT:
            assert EXPECT[name](outs), f'{name}: reference output differs from the surveyed KAT: {outs}'
        doc = {'scenario': scn, 'outcomes': outs, 'final_state': ref_compare.run_reference.last_final_state,
               'generator': 'tests/golden/make_golden.py', 'reference': 'Viasat/nhd v0.3.39, CPython ' + sys.version.split()[0]}
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(doc, f, separators=(',', ':'))
        st = [o['status'] for o in outs]
        print(f'{name}: {len(outs)} pods, placed {st.count("placed")}, none {st.count("none")}, '
              f'other {len(st) - st.count("placed") - st.count("none")}')


if __name__ == '__main__':
    main()
