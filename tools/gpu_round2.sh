#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/quick_bench.py 2 3 4 5 > gpurun_out/quick.log 2>&1
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_chk.so timeout 300 python tools/chk_run.py 4 0 > gpurun_out/chk.log 2>&1
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_prof.so timeout 300 python tools/phase_profile.py 4 > gpurun_out/phase.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1
timeout 900 python tools/scheduler_bench.py 65536 4096 > gpurun_out/scheduler_bench.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
cat gpurun_out/quick.log gpurun_out/chk.log gpurun_out/phase.log; tail -n 3 gpurun_out/pytest_parity.log; tail -n 1 gpurun_out/scheduler_bench.log | cut -c1-1600; cut -c1-3200 gpurun_out/bench.log
