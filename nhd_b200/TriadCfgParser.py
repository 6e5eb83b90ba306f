"""``TriadCfgParser`` — the request codec: a pod's libconfig text <-> ``CfgTopology``.

Host-side mirror of ``nhd/TriadCfgParser.py`` with the reference's class name, constructor and method
names, so ``NHDScheduler.GetCfgParser`` (``NHDScheduler.py:226-232``) can return it unchanged:

* ``CfgToTopology(parseNet)`` (``:337-380``): ``TopologyCfg`` (CPU arch, top-level ``ext_cores``, control
  VLAN name, ``map_type``, ``mod_defs``) + the module instances it points at -> processing groups with
  helper cores, data-path RX/TX core pairs with their Gb/s, CPU workers, GPUs with feeder cores
  (``ParseModGroups`` ``:134-304``), hugepages (``:101-111``) and, for already deployed pods, the
  ``Network_Config`` MACs (``ParseNet`` ``:306-335``).  Every requested core / GPU / VLAN remembers the
  *path* of the setting it came from.
* ``TopologyToCfg()`` (``:413-459``): after placement the physical ids are written back to those paths, the
  GPU map is rebuilt as one nested list, ``Network_Config`` is regenerated (``PopulateNetCfg`` ``:462-497``)
  and the text is re-serialised.  ``TopologyToGpuMap()`` (``:397-410``) gives the ``nvidiaN`` annotation.

Return conventions are the reference's: ``None`` / ``False`` where it returns them, and the same
exceptions escape where it lets them escape (e.g. a missing ``gpu_map`` in a data-path group,
``:231``).  libconfig reading / writing and path addressing come from ``nhd_b200.libconfig`` (the
third-party ``libconf`` / ``magicattr`` the reference imports are not available here).
"""
import collections
import functools

from nhd_b200 import libconfig
from nhd_b200.CfgTopology import (CfgTopology, Core, GPU, NICCoreDirection, NUMASetting, ProcGroup, SMTSetting,
                                  VLANInfo)

_NONE = NICCoreDirection.NIC_CORE_DIRECTION_NONE
_RX = NICCoreDirection.NIC_CORE_DIRECTION_RX
_TX = NICCoreDirection.NIC_CORE_DIRECTION_TX
_GROUP = NUMASetting.LOGICAL_NUMA_GROUP


def _smt(flag):
    return SMTSetting.SMT_ENABLED if flag else SMTSetting.SMT_DISABLED


class TriadCfgParser:
    def __init__(self, dat, isFile, logger=None):
        self.logger = logger
        self.cfg = None
        self.top = CfgTopology()
        if not isFile:
            self.LoadCfgStr(dat)
        else:
            self.LoadCfgFile(dat)

    def _err(self, msg):
        if self.logger is not None:
            self.logger.error(msg)

    # ---- loading (:36-46) -------------------------------------------------------------------
    def LoadCfgStr(self, dat):
        self.cfg = libconfig.loads(dat)

    def LoadCfgFile(self, dat):
        self.cfg = libconfig.load(dat)

    def _get(self, path):
        return libconfig.path_get(self.cfg, path)

    def _core(self, path, speed=0, direction=_NONE, numa=_GROUP):
        return Core(path, speed, direction, numa, int(self._get(path)))

    # ---- reading ------------------------------------------------------------------------------
    def CheckMandatoryFields(self):                          # :49-70
        return all(f in self.cfg.TopologyCfg for f in ('cpu_arch', 'ext_cores', 'kni_vlan'))

    def GetValFromAttrName(self, obj, attr, *args):          # :72-78
        return functools.reduce(lambda o, a: getattr(o, a, *args), [obj] + attr.split('.'))

    def ParseKniDataVlan(self):                              # :80-91
        if 'kni_vlan' not in self.cfg.TopologyCfg:
            return False
        self.top.SetCtrlVlan(VLANInfo(self.cfg.TopologyCfg.kni_vlan, 0))
        return True

    def ParseHugePages(self):                                # :93-104
        if 'Hugepages_GB' not in self.cfg:
            return False
        self.top.hugepages_gb = int(self.cfg.Hugepages_GB)
        return True

    def ParseMiscCores(self):                                # :106-132
        tc = self.cfg.TopologyCfg
        if 'ext_cores' not in tc or 'ext_cores_smt' not in tc:
            return False
        self.top.SetMiscCoreSmt(_smt(tc.ext_cores_smt))
        for path in tc.ext_cores:
            try:
                self.top.AddMiscCore(self._core(path, numa=NUMASetting.LOGICAL_NUMA_DONT_CARE))
            except AttributeError as e:
                self._err(f'Failed to parse field {path} from config file:\n    {e}')
                return False
        return True

    def _parse_dp_group(self, md, mattr, pg) -> bool:        # :184-257
        base = f'{mattr}.{md.dp_group.name}'
        try:
            attr = self._get(base)
        except Exception:
            return False
        if len(attr) != 1:                                   # one NUMA node per data-path group
            return False
        dp = attr[0]
        # the reference's chained comparison (:200): true only when each neighbour pair differs, and
        # evaluated left to right, so a missing later setting may go unnoticed here
        if len(dp.rx_cores) != len(dp.tx_cores) != len(dp.rx_speeds) != len(dp.tx_speeds):
            return False
        pg.SetProcSmt(_smt(md.dp_group.proc_cores_smt))
        try:
            for i in range(len(dp.rx_cores)):
                rx = self._core(f'{base}[0].rx_cores[{i}]', self._get(f'{base}[0].rx_speeds[{i}]'), _RX)
                pg.AddGroupCore(rx)
                tx = self._core(f'{base}[0].tx_cores[{i}]', self._get(f'{base}[0].tx_speeds[{i}]'), _TX)
                pg.AddGroupCore(tx)
                self.top.AddNicPairing(rx, tx)
        except Exception:
            return False
        try:
            for i in range(len(dp.cpu_workers)):
                pg.AddGroupCore(self._core(f'{base}[0].cpu_workers[{i}]'))
        except Exception:
            pass                                             # no CPU workers in this group
        by_device = collections.defaultdict(list)            # placeholder device id -> [(device path, core path)]
        for i in range(len(dp.gpu_map)):                     # a missing gpu_map escapes, as in the reference
            if len(dp.gpu_map[i]) == 2:
                by_device[dp.gpu_map[i][1]].append((f'{base}[0].gpu_map[{i}][1]', f'{base}[0].gpu_map[{i}][0]'))
        gputype = pg.GetGpuType(md.dp_group.gpu_type if 'gpu_type' in md.dp_group else 'ANY')
        for dev, entries in by_device.items():
            feeders = [self._core(core_path) for _, core_path in entries]
            pg.SetGpuType(gputype)
            pg.AddGroupGPU(GPU(feeders, [dev_path for dev_path, _ in entries], gputype, dev))
        return True

    def _parse_nic_cores(self, md, mattr, pg) -> bool:       # :259-297
        spec = md.nic_cores
        if len(spec) != 5:
            return False
        try:
            rx_cores, rx_speeds, tx_cores, tx_speeds = (self._get(f'{mattr}.{spec[k]}') for k in range(4))
        except Exception:
            return False
        if len(rx_cores) != len(rx_speeds) != len(tx_cores) != len(tx_speeds):      # chained, as at :273
            return False
        pg.SetProcSmt(_smt(spec[4]))
        for i in range(len(rx_cores)):
            rx = self._core(f'{mattr}.{spec[0]}[{i}]', self._get(f'{mattr}.{spec[1]}[{i}]'), _RX)
            pg.AddGroupCore(rx)
            tx = self._core(f'{mattr}.{spec[2]}[{i}]', self._get(f'{mattr}.{spec[3]}[{i}]'), _TX)
            pg.AddGroupCore(tx)
            self.top.AddNicPairing(rx, tx)
        return True

    def ParseModGroups(self) -> bool:                        # :134-304
        tc = self.cfg.TopologyCfg
        if 'mod_defs' not in tc or 'map_type' not in tc:
            return False
        self.top.SetTopMapType(tc.map_type)
        for md in tc.mod_defs:
            if md.module not in self.cfg:
                return False
            for idx, mi in enumerate(self.cfg[md.module]):
                pg = ProcGroup()
                mattr = f'{md.module}[{idx}]'
                if 'helper_cores' in md:
                    if 'helper_cores_smt' not in md:
                        return False
                    pg.SetHelperSmt(_smt(md.helper_cores_smt))
                    for hc in md.helper_cores:
                        path = f'{mattr}.{hc}'
                        held = self._get(path)
                        if type(held) == list:               # an array of helper cores under one name
                            for k in range(len(held)):
                                pg.AddMiscCore(self._core(f'{path}[{k}]'))
                        else:
                            pg.AddMiscCore(self._core(path))
                if 'data_vlan' in md:
                    pg.SetDataVlan(VLANInfo(f'{mattr}.{md.data_vlan}', 0))
                if 'dp_group' in md and not self._parse_dp_group(md, mattr, pg):
                    return False
                if 'nic_cores' in md and not self._parse_nic_cores(md, mattr, pg):
                    return False
                self.top.proc_groups.append(pg)
        return True

    def ParseNet(self):                                      # :306-335
        if 'Network_Config' not in self.cfg:
            return False
        for net in self.cfg.Network_Config:
            for i, _ in enumerate(net.rxCores):
                ng = self.top.GetNICGroupFromCoreNumbers(int(net.rxCores[i]), int(net.txCores[i]))
                if ng is None:
                    return False
                ng.AddInterface(net.mac)
                try:
                    ng.SetRxRingSize(int(net.rx_mbufs[i]))
                except AttributeError:
                    pass                                     # configs written before rx_mbufs existed
        return True

    def CfgToTopology(self, parseNet: bool):                 # :337-380
        if 'TopologyCfg' not in self.cfg:
            return None
        if not self.CheckMandatoryFields():
            return None
        if self.top.SetCpuArch(self.cfg.TopologyCfg.cpu_arch) is None:
            return None
        for step in (self.ParseMiscCores, self.ParseKniDataVlan, self.ParseModGroups, self.ParseHugePages):
            if not step():
                return None
        if parseNet and not self.ParseNet():
            return None
        return self.top

    # ---- writing -------------------------------------------------------------------------------
    def SetLibConfigValue(self, name, value):                # :382-395
        if name[-1] == ']':
            libconfig.path_set(self.cfg, name, value)
            return
        pos = name.rfind('.')
        if pos >= 0:
            libconfig.path_get(self.cfg, name[:pos])[name[pos + 1:]] = value
        else:
            self.cfg[name] = value

    def TopologyToGpuMap(self):                              # :397-410
        gpu_annotations = {}
        for pg in self.top.proc_groups:
            index = 0                                        # restarts per group: later groups overwrite
            for g in pg.group_gpus:
                for _ in g.dev_id_names:
                    gpu_annotations['nvidia' + str(index)] = g.device_id
                    index += 1
        return gpu_annotations

    def TopologyToCfg(self) -> str:                          # :413-459
        for c in self.top.misc_cores:
            self.SetLibConfigValue(c.name, c.core)
        self.SetLibConfigValue(self.top.ctrl_vlan.name, self.top.ctrl_vlan.vlan)
        for pg in self.top.proc_groups:
            self.SetLibConfigValue(pg.vlan.name, pg.vlan.vlan)
            for c in pg.proc_cores:
                self.SetLibConfigValue(c.name, c.core)
            for c in pg.misc_cores:
                self.SetLibConfigValue(c.name, c.core)
            if len(pg.group_gpus) > 0:
                gpu_map = []
                for g in pg.group_gpus:
                    for k in range(len(g.dev_id_names)):
                        gpu_map.append((g.cpu_cores[k].core, g.device_id))
                first = pg.group_gpus[0].dev_id_names[0]     # "Mod[0].dp[0].gpu_map[i][1]" -> "Mod[0].dp[0]"
                holder = libconfig.path_get(self.cfg, first[:first.rfind('.')])
                holder['gpu_map'] = tuple(gpu_map)
        self.SetLibConfigValue('Network_Config', self.PopulateNetCfg())
        return libconfig.dumps(self.cfg)

    def PopulateNetCfg(self):                                # :462-497
        by_mac = collections.defaultdict(list)
        for pair in self.top.nic_core_pairing:
            by_mac[pair.mac].append((pair.rx_core.core, pair.tx_core.core, pair.rx_ring_size))
        netconf = []
        ifcnt = 0
        for mac, rows in by_mac.items():
            rxcores, txcores, rings = zip(*rows)
            ips = [f'10.0.0.{x + ifcnt}' for x in range(len(rxcores))]
            netconf.append({'module': f'fake_{ifcnt}', 'ifname': f'fake_if_{ifcnt}', 'mac': mac,
                            'rxCores': list(rxcores), 'txCores': list(txcores), 'rx_mbufs': list(rings),
                            'gwIps': [self.top.data_default_gw] * len(rxcores), 'txIps': ips, 'rxIps': ips,
                            'ts_group': True})
            ifcnt += len(rxcores)
        return tuple(netconf)
