"""Freeze request-codec cases from the UNMODIFIED reference ``nhd.TriadCfgParser`` (build container
only; it runs on the libconf / magicattr stand-ins under oracle/_shim) into
tests/golden/codec/cases.json: config text -> parsed topology (or None / the exception the reference
lets escape) -> rewritten config text and GPU-map annotation after a deterministic fake placement ->
the rewritten text read back with its network section.

    python tests/golden/make_codec_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_sched_loader           # noqa: E402
from tests import triad_cfg as T              # noqa: E402
from tests.test_codec import broken_variants, run_case   # noqa: E402


def main():
    Ref = ref_sched_loader.load_codec().codec.TriadCfgParser
    rng = np.random.default_rng(20260923)
    cases = []

    def add(text, parse_net, seed, note):
        exp = run_case(Ref, text, parse_net, seed)
        cases.append({'note': note, 'seed': seed, 'parse_net': parse_net, 'text': text, 'expected': exp})
        return exp

    for seed in range(48):
        pod = T.codec_pod(rng, 'wild' if seed % 2 else 'mixed')
        text = T.pod_to_cfg(pod, rng, gpu_type=[None, 'V100', 'bogus', '2080Ti'][seed % 4])
        exp = add(text, False, seed, f'generated {seed}')
        add(exp['cfg'], True, seed, f'rewritten {seed} with Network_Config')
        if seed % 6 == 0:
            add(exp['cfg'].replace('rx_mbufs', 'rx_bufs'), True, seed, f'rewritten {seed}, old config without rx_mbufs')
            add(text, True, seed, f'generated {seed}, parseNet without Network_Config')
            for k, bad in enumerate(broken_variants(text, rng)):
                add(bad, False, seed, f'broken {seed}.{k}')
    out = os.path.join(ROOT, 'tests', 'golden', 'codec', 'cases.json')
    with open(out, 'w') as f:
        json.dump(cases, f, indent=0)
    kinds = {}
    for c in cases:
        e = c['expected']
        k = e.get('raises') or ('None' if e['topology'] is None else 'topology')
        kinds[k] = kinds.get(k, 0) + 1
    print(len(cases), 'cases', kinds, os.path.getsize(out) // 1024, 'KiB')


if __name__ == '__main__':
    main()
