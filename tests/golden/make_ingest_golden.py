"""Freezes label-ingest golden vectors from the UNMODIFIED reference (run in the build container):

    python tests/golden/make_ingest_golden.py

For every case: the label dictionary, what the reference's Node.ParseLabels did with it ('ok' / 'false' /
'raises'), and for accepted nodes the nhd_node_rec that nhd_b200.packing.pack_node makes of the REFERENCE's
Node object (hex) or 'unsupported'.  Cases = the nodes of seeded random scenarios plus hand-written edge
cases and seeded single-label mutations of valid dictionaries.  Output: tests/golden/ingest/cases.json.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader                                   # noqa: E402
from nhd_b200 import packing                                     # noqa: E402
from tests import scenarios                                      # noqa: E402

NFD = scenarios.NFD


def reference_outcome(ref, case, layout):
    n = ref.node.Node('x', case['active'])
    try:
        ok = n.ParseLabels(case['labels'])
    except Exception as e:                                       # noqa: BLE001 - the point is "it raises"
        return {'parse': 'raises', 'exc': type(e).__name__}
    if not ok:
        return {'parse': 'false'}
    n.SetHugepages(case['hp_alloc'], case['hp_free'])
    try:
        rec = packing.pack_node(n, layout)
    except packing.UnsupportedError:
        return {'parse': 'ok', 'record': 'unsupported'}
    return {'parse': 'ok', 'record': rec.tobytes().hex(), 'data_vlan': n.data_vlan,
            'res_hugepages_gb': n.mem.res_hugepages_gb, 'n_reserved_cores': len(n.reserved_cores),
            'gpu_device_id': [g.device_id for g in n.gpus], 'nic_ifname': [x.ifname for x in n.nics],
            'groups': list(n.groups)}


def edge_cases():
    mk = scenarios.make_node
    base = mk('e', gpus=[(0, 0, 0x10), (1, 1, 0x20)], nics=[('eth0', 100000, 0, 0x10), ('eth1', 25000, 1, 0x20)])
    out = []

    def add(name, labels, **kw):
        c = {'name': name, 'labels': labels, 'active': kw.get('active', True), 'hp_alloc': kw.get('hp_alloc', 64),
             'hp_free': kw.get('hp_free', 60)}
        out.append(c)

    def without(key):
        return {k: v for k, v in base['labels'].items() if k != key}

    def with_(**kv):
        d = dict(base['labels'])
        d.update(kv)
        return d

    add('base', base['labels'])
    add('inactive', base['labels'], active=False)
    add('no_num_cores', without(NFD + 'nfd-extras-cpu.num_cores'))
    add('no_sockets', without(NFD + 'nfd-extras-cpu.numSockets'))
    add('no_vlan', without('DATA_PLANE_VLAN'))
    add('no_gw', without('DATA_DEFAULT_GW'))
    add('no_isolcpus', without(NFD + 'nfd-extras-cpu.isolcpus'))
    add('no_smt', without(NFD + 'cpu-hardware_multithreading'))
    add('res_hugepages', with_(RES_HUGEPAGES_GB='6'))
    add('bad_res_hugepages', with_(RES_HUGEPAGES_GB='six'))
    add('bad_vlan', with_(DATA_PLANE_VLAN='v100'))
    add('spaces_in_ints', with_(**{NFD + 'nfd-extras-cpu.num_cores': ' 32 ', 'DATA_PLANE_VLAN': '+7'}))
    add('underscore_int', with_(**{NFD + 'nfd-extras-cpu.num_cores': '3_2'}))
    add('zero_sockets', with_(**{NFD + 'nfd-extras-cpu.numSockets': '0'}))
    add('three_sockets_irregular', with_(**{NFD + 'nfd-extras-cpu.numSockets': '3'}))
    add('five_sockets', with_(**{NFD + 'nfd-extras-cpu.numSockets': '5', NFD + 'nfd-extras-cpu.num_cores': '40'}))
    add('four_sockets', mk('f', sockets=4, phys_cores=64, nics=[('a', 100000, 3, 1), ('b', 100000, 0, 2)])['labels'])
    add('one_socket', mk('g', sockets=1, phys_cores=16, smt=False, nics=[('a', 100000, 0, 1)])['labels'])
    add('too_many_cores', mk('h', sockets=2, phys_cores=160, smt=True)['labels'])
    add('max_cores', mk('h2', sockets=2, phys_cores=128, smt=True)['labels'])
    add('isol_triple_dash', with_(**{NFD + 'nfd-extras-cpu.isolcpus': '2-5-9_20-31'}))
    add('isol_reversed', with_(**{NFD + 'nfd-extras-cpu.isolcpus': '9-2_4'}))
    add('isol_beyond', with_(**{NFD + 'nfd-extras-cpu.isolcpus': '2-500'}))
    add('isol_empty_part', with_(**{NFD + 'nfd-extras-cpu.isolcpus': '2-5,,7'}))
    add('isol_garbage', with_(**{NFD + 'nfd-extras-cpu.isolcpus': '2-x'}))
    add('maint_draining', with_(**{'sigproc.viasat.io/maintenance': 'Draining'}))
    add('maint_not_scheduled', with_(**{'sigproc.viasat.io/maintenance': 'NOT_SCHEDULED'}))
    add('groups', with_(NHD_GROUP='edge.default.gpu-a'))
    add('groups_empty_name', with_(NHD_GROUP='a..b'))
    add('nic_down', mk('i', nics=[('eth0', 0, 0, 1), ('eth1', 100000, 1, 2)])['labels'])
    add('nic_slow', mk('j', nics=[('eth0', 10000, 0, 1), ('eth1', 11000, 1, 2)])['labels'])
    add('sriov_pf_excluded', mk('k', sriov_pfs=['pf0'], nics=[('pf0', 100000, 0, 1), ('vf0', 25000, 0, 1),
                                                                ('vf1', 25000, 1, 2)])['labels'])
    add('nic_short_key', with_(**{NFD + 'nfd-extras-nic.eth9.mlx.aabbccddeeff.100000Mbs.0.1': 'true'}))
    add('nic_bad_numa', with_(**{NFD + 'nfd-extras-nic.eth9.mlx.aabbccddeeff.100000Mbs.z.1.1.0': 'true'}))
    add('nic_numa_out_of_range', with_(**{NFD + 'nfd-extras-nic.eth9.mlx.aabbccddeeff.100000Mbs.2.1.1.0': 'true'}))
    add('nic_hex_prefix', with_(**{NFD + 'nfd-extras-nic.eth9.mlx.aabbccddeeff.100000Mbs.1.0x3f.0x1.0': 'true'}))
    add('nic_down_bad_port', with_(**{NFD + 'nfd-extras-nic.eth9.mlx.aabbccddeeff.down.1.1.1.p': 'true'}))
    add('nic_speed_garbage', with_(**{NFD + 'nfd-extras-nic.eth9.mlx.aabbccddeeff.fastMbs.1.1.1.0': 'true'}))
    add('gpu_short_key', with_(**{NFD + 'nfd-extras-gpu.4.V100.1': 'true'}))
    add('gpu_numa_out_of_range', with_(**{NFD + 'nfd-extras-gpu.4.V100.3.1f': 'true'}))
    add('gpu_bad_device', with_(**{NFD + 'nfd-extras-gpu.four.V100.1.1f': 'true'}))
    add('seventeen_gpus', mk('l', gpus=[(i, i % 2, i // 2) for i in range(17)])['labels'])
    add('sixteen_gpus', mk('l2', gpus=[(i, i % 2, i // 2) for i in range(16)])['labels'])
    add('seventeen_switches', mk('m', gpus=[(i, i % 2, 0x40 + i) for i in range(9)],
                                 nics=[(f'e{i}', 100000, i % 2, 0x80 + i) for i in range(8)])['labels'])
    add('thirty_three_nics', mk('n', nics=[(f'e{i}', 100000, i % 2, 1 + i % 2) for i in range(33)])['labels'])
    add('thirty_two_nics', mk('n2', nics=[(f'e{i}', 25000, i % 2, 1 + i % 2) for i in range(32)])['labels'])
    add('seventeen_speeds_first', mk('o', nics=[(f'e{i}', 11000 + 1000 * i, i % 2, 1) for i in range(9)])['labels'])
    add('seventeen_speeds_second', mk('p', nics=[(f'e{i}', 41000 + 1000 * i, i % 2, 1) for i in range(9)])['labels'])
    add('no_labels_at_all', {})
    return out


def mutations(rng, n):
    """single-label damage to valid dictionaries"""
    out = []
    for i in range(n):
        scn = scenarios.random_scenario(91000 + i, n_nodes=1, n_pods=1, flavor=rng.choice(['mixed', 'wild', 'vf', 'big']))
        nd = scn['nodes'][0]
        labels = dict(nd['labels'])
        keys = list(labels.keys())
        k = rng.choice(keys)
        how = rng.choice(['drop', 'value', 'field', 'dup'])
        if how == 'drop':
            del labels[k]
        elif how == 'value':
            labels[k] = rng.choice(['', '0', '-1', 'x', '1e3', ' 12', '0x10', '7_7', '2-', '-2', '1,2', '3-1_0-0'])
        elif how == 'field' and '.' in k:
            parts = k.split('.')
            j = rng.randrange(len(parts))
            parts[j] = rng.choice(['', '0', 'zz', '-1', '9', '0x2', 'ff', '40000Mbs', 'Mbs', '1Mbs0'])
            v = labels.pop(k)
            labels = {('.'.join(parts) if kk is None else kk): vv for kk, vv in
                      [(kk2, vv2) for kk2, vv2 in labels.items()] + [(None, v)]}
        else:
            labels[k + '.extra'] = labels[k]
        out.append({'name': f'mut{i}', 'labels': labels, 'active': nd.get('active', True),
                    'hp_alloc': nd['hp_alloc'], 'hp_free': nd['hp_free']})
    return out


def main():
    ref = ref_loader.load()
    rng = random.Random(20260923)
    cases = []
    for seed, flavor in ((7001, 'mixed'), (7002, 'wild'), (7003, 'vf'), (7004, 'big')):
        scn = scenarios.random_scenario(seed, n_nodes=12, n_pods=1, flavor=flavor)
        for nd in scn['nodes']:
            cases.append({'name': f'{flavor}:{nd["name"]}', 'labels': nd['labels'], 'active': nd.get('active', True),
                          'hp_alloc': nd['hp_alloc'], 'hp_free': nd['hp_free']})
    cases += edge_cases()
    cases += mutations(rng, 120)
    layout = packing.ClusterLayout()        # one cluster dictionary across all cases, in case order
    for c in cases:
        c['expect'] = reference_outcome(ref, c, layout)
    out = {'cases': cases, 'speed_table': [float(x) for x in layout.speed_table()],
           'group_bits': layout.group_bits}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ingest', 'cases.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0, sort_keys=False)
    kinds = {}
    for c in cases:
        k = c['expect']['parse'] + ('/' + ('unsupported' if c['expect'].get('record') == 'unsupported' else 'rec')
                                    if c['expect']['parse'] == 'ok' else '')
        kinds[k] = kinds.get(k, 0) + 1
    print(len(cases), 'cases', kinds, '->', path)


if __name__ == '__main__':
    main()
