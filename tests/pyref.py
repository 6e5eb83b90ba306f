"""Tiny Python restatements used as ground truth by tests: they run the reference's own
expressions on REAL CPython sets, so any set-order model can be checked against the
interpreter that is running the test."""
import itertools


def choose_mapping(K, G, maskA, maskB, maskC):
    """Matcher.py:113-141,175-220,344-368,428-444 on real sets.  masks are sets of tuple indices
    (product order).  Returns (gtuple index, misc numa) or None."""
    prodG = list(itertools.product(range(K), repeat=G))
    prodB = list(itertools.product(range(K), repeat=G + 1))
    stmp = set()
    for i, p in enumerate(prodG):
        if i in maskA:
            stmp.add(p)
    a_list = list(stmp)
    stmp = set()
    for i, p in enumerate(prodB):
        if i in maskB:
            stmp.add(p)
    b_list = list(stmp)
    nic_tuples = [p for i, p in enumerate(prodG) if i in maskC]
    if not a_list or not b_list or not nic_tuples:
        return None
    gpu_tuples = [x for x in a_list]
    cpu_tuples = [x[:-1] for x in b_list]
    intersect = list(set(gpu_tuples) & set(cpu_tuples) & set(nic_tuples))
    if len(intersect) == 0:
        return None
    gl = a_list
    diff = set(a_list) - set(intersect)
    if len(diff):
        gl = intersect

    def node_delta(x):
        el = [gl[x].count(y) for y in range(K)]
        return max(el) - min(el)
    gidx, gval = 0, node_delta(0)
    for t in range(1, len(gl)):
        tmp = node_delta(t)
        if tmp > gval:
            gidx, gval = t, tmp
    gtuple = gl[gidx]
    cabbr = [x[:-1] for x in b_list]
    ctuple = b_list[cabbr.index(gtuple)]
    return prodG.index(gtuple), ctuple[-1]
