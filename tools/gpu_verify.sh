#!/bin/bash
# Developer script: the heavy hardware checks — full-size parity, whole -m gpu suite, compute-sanitizer.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/pytest_full_size.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py > gpurun_out/pytest_gpu.log 2>&1
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_run.py > gpurun_out/sanitizer_$tool.log 2>&1
done
tail -n 6 gpurun_out/pytest_full_size.log gpurun_out/pytest_gpu.log; tail -n 8 gpurun_out/sanitizer_*.log
