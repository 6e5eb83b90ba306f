#!/bin/bash
# Developer script for `gpurun --gpus N`: the bench at N, …, 2 and 1 ranks (largest first: the GPU budget may cut the
# visit short), then multi-rank parity against the oracle.
N=${1:-2}
mkdir -p gpurun_out
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n > gpurun_out/bench_n$n.log 2>&1
  fi
done
timeout 300 python bench.py --no-extra > gpurun_out/bench_n1.log 2>&1
for f in gpurun_out/bench_n*.log; do tail -n 1 $f | cut -c1-700; echo; done
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_multirank_$N.log 2>&1
tail -n 5 gpurun_out/pytest_multirank_$N.log
