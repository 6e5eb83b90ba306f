#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/quick_bench.py 2 3 4 > gpurun_out/quick.log 2>&1
NHD_B200_LIB=$PWD/nhd_b200/libnhd_b200_chk.so timeout 300 python tools/chk_run.py 4 0 > gpurun_out/chk.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
cat gpurun_out/quick.log gpurun_out/chk.log; tail -n 3 gpurun_out/pytest_parity.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'])
print(json.dumps(d.get('extra'), indent=1))
PY
