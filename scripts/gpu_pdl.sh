#!/bin/bash
# PDL / fused-head experiment: correctness first, then the default bench under each knob combination.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_net.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_net.log
tail -8 gpurun_out/pytest_net.log
for pdl in 0 1 2; do for head in 0 1; do
  OKL_PDL=$pdl OKL_FUSE_HEAD=$head timeout 300 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/b_${pdl}_${head}.log 2> gpurun_out/b_${pdl}_${head}.err
  echo "pdl=$pdl head=$head exit $?"
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/b_${pdl}_${head}.log').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','final_loss')}, 'e2e', l['e2e'] and l['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/b_${pdl}_${head}.err').read()[-1500:])
PY
done; done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
