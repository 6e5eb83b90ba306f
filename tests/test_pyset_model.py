"""Pins both CPython-set models (the oracle's pyset_model.c and the product's nhd_core.cuh
emulator) against the interpreter running the tests (reference needs CPython >= 3.8)."""
import ctypes
import itertools
import random

import numpy as np
import pytest

from tests import pyref


def _oracle_set_list(lib, tuples):
    """list(set built by adding `tuples` in order) according to oracle/pyset_model.c"""
    L = lib.lib()
    class PySet(ctypes.Structure):
        _fields_ = [('slots', ctypes.c_byte * (24 * 4096)), ('mask', ctypes.c_size_t),
                    ('fill', ctypes.c_size_t), ('used', ctypes.c_size_t)]
    s = PySet()
    L.pyset_init(ctypes.byref(s))
    L.pyset_tuple_key.restype = ctypes.c_uint64
    L.py_hash_tuple.restype = ctypes.c_uint64
    L.pyset_add.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
    for t in tuples:
        arr = (ctypes.c_int * len(t))(*t)
        L.pyset_add(ctypes.byref(s), L.pyset_tuple_key(arr, len(t)), L.py_hash_tuple(arr, len(t)))
    keys = (ctypes.c_uint64 * 4096)()
    L.pyset_list.restype = ctypes.c_size_t
    n = L.pyset_list(ctypes.byref(s), keys, None)
    out = []
    for i in range(n):
        k = keys[i]
        ln = k >> 56
        out.append(tuple((k >> (4 * j)) & 0xF for j in range(ln)))
    return out


def test_tuple_hash_matches_interpreter(oracle_lib, emu):
    L = oracle_lib.lib()
    L.py_hash_tuple.restype = ctypes.c_uint64
    for K in (1, 2, 3, 4):
        for Ln in range(0, 6):
            for idx, t in enumerate(itertools.product(range(K), repeat=Ln)):
                want = hash(t) & 0xFFFFFFFFFFFFFFFF
                arr = (ctypes.c_int * max(Ln, 1))(*t)
                assert L.py_hash_tuple(arr, Ln) == want
                if K ** Ln <= 256 and Ln >= 1:
                    assert emu.nhd_emu_tuple_hash(idx, K, Ln) == want
    assert hash(()) == 5740354900026072187
    assert hash((0,)) == -8753497827991233192


def test_known_set_orders():
    # SURVEY.md appendix D set-order KATs (CPython 3.8 - 3.12)
    full = lambda L: list({p for p in itertools.product(range(2), repeat=L)})
    s = set()
    for p in itertools.product(range(2), repeat=2):
        s.add(p)
    assert list(s) == [(0, 1), (1, 0), (1, 1), (0, 0)]
    s = set()
    for p in itertools.product(range(2), repeat=3):
        s.add(p)
    assert list(s) == [(1, 0, 1), (1, 1, 0), (0, 1, 0), (0, 0, 0), (1, 0, 0), (0, 0, 1), (1, 1, 1), (0, 1, 1)]


@pytest.mark.parametrize('K,L', [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 2), (3, 3), (4, 2), (4, 4), (3, 4)])
def test_oracle_set_order_random_subsets(oracle_lib, K, L):
    rng = random.Random(K * 100 + L)
    prod = list(itertools.product(range(K), repeat=L))
    for trial in range(150):
        k = rng.randint(0, len(prod))
        sub = [p for p in prod if rng.random() < k / max(1, len(prod))]
        s = set()
        for p in sub:
            s.add(p)
        assert _oracle_set_list(oracle_lib, sub) == list(s)
        # arbitrary insertion order as well (set(A_list) is built in list order)
        rng.shuffle(sub)
        s = set()
        for p in sub:
            s.add(p)
        assert _oracle_set_list(oracle_lib, sub) == list(s)


def _mask_words(idxs):
    w = [0, 0, 0, 0]
    for i in idxs:
        w[i >> 6] |= 1 << (i & 63)
    return (ctypes.c_uint64 * 4)(*w)


@pytest.mark.parametrize('K,G', [(1, 1), (1, 3), (2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3), (4, 1), (4, 2), (4, 3)])
def test_product_choose_mapping_vs_interpreter(emu, K, G):
    """The product's emulator picks the same (gtuple, misc NUMA) as the reference's
    expressions evaluated on real CPython sets."""
    rng = random.Random(K * 10 + G)
    np_, nq = K ** G, K ** (G + 1)
    n_trials = 4000 if nq <= 32 else 600
    exhaustive = (K == 2 and G <= 2)
    cases = []
    if exhaustive:
        for a in range(1, 1 << np_):
            for c in range(1, 1 << np_):
                for _ in range(6):
                    cases.append((a, rng.randrange(1, 1 << nq), c))
    else:
        for _ in range(n_trials):
            dens = rng.choice([0.2, 0.5, 0.8, 0.95, 1.0])
            pick = lambda n: sum(1 << i for i in range(n) if rng.random() < dens)
            cases.append((pick(np_), pick(nq), pick(np_)))
    p = ctypes.c_int()
    m = ctypes.c_int()
    for a, b, c in cases:
        sa = {i for i in range(np_) if a >> i & 1}
        sb = {i for i in range(nq) if b >> i & 1}
        sc = {i for i in range(np_) if c >> i & 1}
        want = pyref.choose_mapping(K, G, sa, sb, sc)
        got = emu.nhd_emu_choose(K, G, _mask_words(sa), _mask_words(sb), _mask_words(sc),
                                 ctypes.byref(p), ctypes.byref(m))
        if want is None:
            assert got == 0, (K, G, a, b, c)
        else:
            assert got == 1 and (p.value, m.value) == want, (K, G, a, b, c, want, (p.value, m.value))


def test_claimed_nic_order(emu):
    rng = random.Random(5)
    out = (ctypes.c_uint8 * 8)()
    for _ in range(3000):
        n = rng.randint(0, 4)
        li = [rng.randrange(0, 32) for _ in range(n)]
        want = list({x for x in li})
        arr = (ctypes.c_uint8 * 4)(*(li + [0] * (4 - n)))
        k = emu.nhd_emu_claim_order(arr, n, out)
        assert [out[i] for i in range(k)] == want, (li, want)
