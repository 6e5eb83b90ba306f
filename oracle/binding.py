"""ctypes binding of the C oracle (oracle/_ref/libnhd_oracle.so).  ORACLE — tests only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_ref', 'libnhd_oracle.so')


class OracleParams(ctypes.Structure):
    _fields_ = [('nic_bw_avail_percent', ctypes.c_double), ('min_busy_secs', ctypes.c_double)]


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ('nhd_oracle.c', 'pyset_model.c', 'nhd_oracle.h', 'pyset_model.h')]
    srcs.append(os.path.join(_HERE, '..', 'include', 'nhd_b200.h'))
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(['make', '-C', _HERE, '-B'], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.nhd_oracle_solve.restype = ctypes.c_int
        L.nhd_oracle_solve.argtypes = [ctypes.POINTER(OracleParams), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.nhd_oracle_solve_mt.restype = ctypes.c_int
        L.nhd_oracle_solve_mt.argtypes = L.nhd_oracle_solve.argtypes + [ctypes.c_int]
        L.nhd_oracle_candidates.restype = ctypes.c_int
        L.nhd_oracle_candidates.argtypes = [ctypes.POINTER(OracleParams), ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
        for name, res, args in (
                ('py_hash_tuple', ctypes.c_uint64, [ctypes.POINTER(ctypes.c_int), ctypes.c_int]),):
            getattr(L, name).restype = res
            getattr(L, name).argtypes = args
        _lib = L
    return _lib


def solve(recs, speed_table, pods, now, nic_bw_avail_percent=0.9, min_busy_secs=30.0, threads=1):
    """Run the oracle over a batch.  Returns (bindings, final_records); inputs are not modified.
    ``threads`` > 1 splits each pod's walk over the nodes over host threads (same result)."""
    from nhd_b200 import wire
    recs = np.ascontiguousarray(recs, dtype=wire.NODE_DTYPE).copy()
    pods = np.ascontiguousarray(pods, dtype=wire.POD_DTYPE)
    now = np.ascontiguousarray(now, dtype='<f8')
    speed = np.ascontiguousarray(speed_table, dtype='<f8')
    assert len(now) == len(pods)
    out = np.zeros(len(pods), dtype=wire.BINDING_DTYPE)
    p = OracleParams(nic_bw_avail_percent, min_busy_secs)
    if threads > 1:
        rc = lib().nhd_oracle_solve_mt(ctypes.byref(p), speed.ctypes.data, len(recs), recs.ctypes.data,
                                       len(pods), pods.ctypes.data, now.ctypes.data, out.ctypes.data, int(threads))
    else:
        rc = lib().nhd_oracle_solve(ctypes.byref(p), speed.ctypes.data, len(recs), recs.ctypes.data,
                                    len(pods), pods.ctypes.data, now.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f'nhd_oracle_solve failed: {rc}')
    return out, recs


def candidates(recs, speed_table, pod, now, nic_bw_avail_percent=0.9, min_busy_secs=30.0):
    """filts[1] membership (after IntersectResources) of every node for one pod."""
    from nhd_b200 import wire
    recs = np.ascontiguousarray(recs, dtype=wire.NODE_DTYPE)
    pod = np.ascontiguousarray(pod, dtype=wire.POD_DTYPE).reshape(1)
    speed = np.ascontiguousarray(speed_table, dtype='<f8')
    out = np.zeros(len(recs), dtype=np.uint8)
    p = OracleParams(nic_bw_avail_percent, min_busy_secs)
    rc = lib().nhd_oracle_candidates(ctypes.byref(p), speed.ctypes.data, len(recs), recs.ctypes.data,
                                     pod.ctypes.data, float(now), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f'nhd_oracle_candidates failed: {rc}')
    return out
