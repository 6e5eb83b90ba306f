#!/bin/bash
# Developer script: one GPU-box visit = quick timings, checks build, phase profile, parity tests, bench.  Output under gpurun_out/.
mkdir -p gpurun_out
timeout 300 python tools/quick_bench.py 2 3 4 > gpurun_out/quick.log 2>&1
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_chk.so timeout 300 python tools/chk_run.py 4 0 > gpurun_out/chk.log 2>&1
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_prof.so timeout 300 python tools/phase_profile.py 4 > gpurun_out/phase.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
tail -n 30 gpurun_out/quick.log gpurun_out/chk.log gpurun_out/phase.log gpurun_out/pytest_parity.log; cut -c1-1500 gpurun_out/bench.log
