"""In-memory stand-ins for the two things ``NHDScheduler`` talks to besides ``Node`` and
``Matcher``: the Kubernetes manager (``nhd/K8SMgr.py``) and the per-pod config parser
(``nhd/TriadCfgParser.py``).  TEST INFRASTRUCTURE: the same objects drive the UNMODIFIED
reference scheduler (build container) and this repo's ``nhd_b200.NHDScheduler``, so the two
can be compared call for call.

``FakeK8s`` keeps the method names, argument order and return conventions of ``K8SMgr``
(``K8SMgr.py:55-230, 284-530``) for every method ``NHDScheduler`` uses, and records what the
scheduler did to each pod (events, annotations, binds).  Failures are injected per pod:
``fail`` is a set of step names out of ``podobj`` (``GetPodObj`` -> None), ``cfg`` (config does
not parse), ``nad``, ``gpumap``, ``annotate``, ``bind`` (the K8s write returns False).

``JsonCfgParser`` plays ``TriadCfgParser``: the pod's config text is JSON
``{'pod': <tests.scenarios pod description>, 'assigned': <extract_result>}`` instead of
libconfig (``libconf`` / ``magicattr`` are not available here); it offers the three
calls the scheduler makes — ``CfgToTopology(parseNet)``, ``TopologyToCfg()``,
``TopologyToGpuMap()`` (``TriadCfgParser.py:337-380, 397-459``).
"""
import json

from tests import scenarios

CFG_ANNOTATION = 'sigproc.viasat.io/nhd_config'          # K8SMgr.py:139
GROUPS_ANNOTATION = 'sigproc.viasat.io/nhd_groups'       # K8SMgr.py:160
GPUMAP_ANNOTATION = 'sigproc.viasat.io/nhd_gpu_map'
NAD_ANNOTATION = 'k8s.v1.cni.cncf.io/networks'           # K8SMgr.py:284-298


def fill_result(top, res, with_net=True):
    """Inverse of ``scenarios.extract_result``: physical ids back into a topology."""
    cores, gpus = list(res['cores']), list(res['gpus'])
    ci = gi = 0
    for pi, pg in enumerate(top.proc_groups):
        for g in pg.group_gpus:
            g.device_id = gpus[gi]
            gi += 1
            for c in g.cpu_cores:
                c.core = cores[ci]
                ci += 1
        for c in list(pg.proc_cores) + list(pg.misc_cores):
            c.core = cores[ci]
            ci += 1
        pg.vlan.vlan = res['vlans'][pi]
    for c in top.misc_cores:
        c.core = cores[ci]
        ci += 1
    top.ctrl_vlan.vlan = res['ctrl_vlan']
    top.data_default_gw = res['gw']
    if with_net:                                          # ParseNet, TriadCfgParser.py:306-335
        for pair, mac in zip(top.nic_core_pairing, res['macs']):
            pair.AddInterface(mac)


class JsonCfgParser:
    def __init__(self, cfgstr, cfg_mod):
        self.cfg_mod = cfg_mod
        self.doc = None
        self.top = None
        try:
            self.doc = json.loads(cfgstr)
        except (TypeError, ValueError):
            pass

    def CfgToTopology(self, parseNet: bool):
        if not self.doc or self.doc.get('broken') or 'pod' not in self.doc:
            return None
        self.top = scenarios.build_top(self.doc['pod'], self.cfg_mod)
        if 'assigned' in self.doc:
            fill_result(self.top, self.doc['assigned'], with_net=parseNet)
        return self.top

    def TopologyToCfg(self) -> str:
        return json.dumps({'pod': self.doc['pod'], 'assigned': scenarios.extract_result(self.top)}, sort_keys=True)

    def TopologyToGpuMap(self):
        out = {}
        for pg in self.top.proc_groups:                   # TriadCfgParser.py:397-410
            index = 0
            for g in pg.group_gpus:
                for _ in g.dev_id_names:
                    out['nvidia' + str(index)] = g.device_id
                    index += 1
        return out


class FakeK8s:
    def __init__(self, scn_nodes, codec='json'):
        self.codec = codec                                # 'json' (JsonCfgParser) or 'triad' (libconfig text)
        self.node_defs = {n['name']: n for n in scn_nodes}
        self.pods = {}                                    # (ns, name) -> dict, insertion order = list order
        self.events = {}
        self.binds = []
        self.calls = 0

    # ---- test-side population ------------------------------------------------------
    def add_pod(self, ns, name, pod_desc, uid=None, fail=(), phase='Pending', broken_cfg=False):
        doc = {'pod': pod_desc}
        if broken_cfg:
            doc['broken'] = True
        if self.codec == 'triad':
            import zlib
            import numpy as np
            from tests import triad_cfg
            cfg_text = triad_cfg.pod_to_cfg(pod_desc, np.random.default_rng(zlib.crc32(f'{ns}/{name}'.encode())))
        else:
            cfg_text = json.dumps(doc, sort_keys=True)
        annotations = {}
        if pod_desc.get('groups') and list(pod_desc['groups']) != ['default']:
            annotations[GROUPS_ANNOTATION] = ','.join(pod_desc['groups'])
        self.pods[(ns, name)] = {'uid': uid or f'uid-{ns}-{name}', 'phase': phase, 'node': None,
                                 'annotations': annotations, 'cfg': cfg_text,
                                 'fail': set(fail),
                                 'requests': {'hugepages-1Gi': f'{pod_desc.get("hugepages", 0)}Gi'}}
        self.events[(ns, name)] = []

    def delete_pod(self, ns, name):
        self.pods.pop((ns, name), None)

    # ---- nodes (K8SMgr.py:55-110, 167-192) ------------------------------------------
    def GetNodes(self):
        return list(self.node_defs)

    def IsNodeActive(self, node):
        return self.node_defs[node].get('active', True)

    def GetNodeAddr(self, name):
        return '10.0.0.' + str(list(self.node_defs).index(name) % 250 + 1)

    def GetNodeLabels(self, name):
        return self.node_defs[name]['labels']

    def GetNodeHugepageResources(self, node):
        d = self.node_defs[node]
        return (d['hp_alloc'], d['hp_free'])

    # ---- pods, read side --------------------------------------------------------------
    def ServicePods(self, sched_name):                    # K8SMgr.py:227-242
        return {(ns, name, p['uid']): (p['phase'], p['node']) for (ns, name), p in self.pods.items()}

    def GetScheduledPods(self, sched_name):               # K8SMgr.py:204-213
        return [(name, ns, p['uid'], p['phase']) for (ns, name), p in self.pods.items()]

    def GetPodObj(self, pod, ns):
        self.calls += 1
        p = self.pods.get((ns, pod))
        if p is None or 'podobj' in p['fail']:
            return None
        return ('podobj', ns, pod)

    def GetCfgMap(self, pod, ns):                         # K8SMgr.py:328-357: (configmap name, text)
        p = self.pods[(ns, pod)]
        if 'cfg' in p['fail']:                            # a config CfgToTopology refuses (returns None)
            return (f'{pod}-cfg', p['cfg'].replace('TopologyCfg', 'TopoCfg') if self.codec == 'triad' else '{not json')
        return (f'{pod}-cfg', p['cfg'])

    def GetCfgType(self, pod, ns):
        return 'triad'

    def GetPodNodeGroups(self, pod, ns):                  # K8SMgr.py:152-165
        p = self.pods.get((ns, pod))
        if p is None or GROUPS_ANNOTATION not in p['annotations']:
            return ['default']
        return p['annotations'][GROUPS_ANNOTATION].split(',')

    def GetPodAnnotations(self, podname, ns):
        p = self.pods.get((ns, podname))
        return None if p is None else p['annotations']

    def GetCfgAnnotations(self, pod, ns):                 # K8SMgr.py:137-150
        annot = self.GetPodAnnotations(pod, ns)
        if annot is None or CFG_ANNOTATION not in annot:
            return False
        return annot[CFG_ANNOTATION]

    def GetPodNode(self, pod, ns):
        p = self.pods.get((ns, pod))
        return '' if p is None or not p['node'] else p['node']

    def GetRequestedPodResources(self, pod, ns):
        p = self.pods.get((ns, pod))
        return {} if p is None else p['requests']

    # ---- pods, write side ---------------------------------------------------------------
    def GeneratePodEvent(self, podobj, podname, ns, reason, _type, message):
        self.events[(ns, podname)].append([reason, _type.name, message])

    def _write(self, ns, pod, step, key, value):
        self.calls += 1
        p = self.pods.get((ns, pod))
        if p is None or step in p['fail']:
            return False
        p['annotations'][key] = value
        return True

    def AddNADToPod(self, pod, ns, nads):
        return self._write(ns, pod, 'nad', NAD_ANNOTATION, nads)

    def AnnotatePodGpuMap(self, ns, podname, gpumap):
        return self._write(ns, podname, 'gpumap', GPUMAP_ANNOTATION, json.dumps(gpumap, sort_keys=True))

    def AnnotatePodConfig(self, ns, podname, configstr):
        return self._write(ns, podname, 'annotate', CFG_ANNOTATION, configstr)

    def BindPodToNode(self, podname, node, ns):           # K8SMgr.py:468-492
        self.calls += 1
        p = self.pods.get((ns, podname))
        if p is None or 'bind' in p['fail']:
            return False
        p['node'] = node
        self.binds.append([ns, podname, node])
        return True

    # ---- what a test compares -------------------------------------------------------------
    def transcript(self):
        return {'events': {f'{ns}/{name}': ev for (ns, name), ev in self.events.items()},
                'binds': self.binds,
                'annotations': {f'{ns}/{name}': dict(p['annotations']) for (ns, name), p in self.pods.items()},
                'pod_nodes': {f'{ns}/{name}': p['node'] for (ns, name), p in self.pods.items()}}
