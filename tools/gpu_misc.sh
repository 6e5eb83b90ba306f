#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_run.py > gpurun_out/sanitizer_racecheck.log 2>&1
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_prof.so timeout 300 python tools/phase_profile.py 4 > gpurun_out/phase.log 2>&1
timeout 900 python tools/scheduler_bench.py 65536 4096 > gpurun_out/scheduler_bench.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
tail -n 5 gpurun_out/sanitizer_racecheck.log; cat gpurun_out/phase.log; tail -n 2 gpurun_out/scheduler_bench.log | cut -c1-1200; cut -c1-3000 gpurun_out/bench.log
