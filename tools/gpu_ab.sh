#!/bin/bash
# one visit to the GPU box: A/B of three library builds, the quick part of the GPU suite, the contract bench,
# the ncu launch list, the full-size parity tests
mkdir -p gpurun_out
: > gpurun_out/ab.txt
# builds of earlier revisions go under nhd_b200/ab/ (nvcc on `git show <rev>:nhd_b200/csrc/...`), named lib_<rev>.so
for L in $(cd nhd_b200 && ls ab/lib_*.so 2>/dev/null) libnhd_b200.so; do
  NHD_B200_LIB=$L timeout 150 python tools/ab_bench.py 4 >> gpurun_out/ab.txt 2>> gpurun_out/ab_err.log
done
cat gpurun_out/ab.txt
timeout 200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_err.log
cut -c1-400 gpurun_out/bench_final.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-extra --cpu-sample-pods 8 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/launches.csv | cut -c1-200
timeout 150 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_full.log 2>&1
tail -3 gpurun_out/pytest_full.log
