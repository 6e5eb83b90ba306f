#!/bin/bash
mkdir -p gpurun_out
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_prof.so timeout 300 python tools/phase_profile.py 4 > gpurun_out/phase.log 2>&1
NHD_SWEEP_DEBUG=2 NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_prof.so timeout 300 python tools/phase_profile.py 4 > gpurun_out/phase_nosplit.log 2>&1
cat gpurun_out/phase.log gpurun_out/phase_nosplit.log
