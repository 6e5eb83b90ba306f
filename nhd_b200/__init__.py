"""nhd_b200 — B200-native batched placement solver for the NHD scheduler's hot path.

Layout:
  csrc/            CUDA kernels (sm_100a) and the C-ABI library (include/nhd_b200.h)
  _lib, solver     ctypes binding and handle wrapper
  wire, packing    packed record formats and object <-> record conversion
  CfgTopology, Node, Matcher   host-side mirrors of the reference interface
"""
__all__ = ['CfgTopology', 'Node', 'Matcher', 'packing', 'wire', 'solver']
