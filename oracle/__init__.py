"""ORACLE package — test infrastructure only.

Holds the CPU restatement of the reference hot path (nhd_oracle.c + pyset_model.c),
its ctypes binding (binding.py) and the loader for the unmodified reference modules
(ref_loader.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; nhd_b200 (the product) never does.
"""
