#!/bin/bash
# run with: gpurun --gpus N -- 'bash scripts/gpu_dp2.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp.py -m gpu -x -q 2>&1 | tail -5
for ee in ${EES:-1 0}; do
  OKL_EARLY_EXCHANGE=$ee timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$ee bench.py --gpus $N --steps 1000 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/dp_${N}_${ee}.log 2> gpurun_out/dp_${N}_${ee}.err
  echo "early_exchange=$ee exit $?"
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/dp_${N}_${ee}.log').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', l['e2e'] and (l['e2e']['value'], l['e2e']['ms_per_step']))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/dp_${N}_${ee}.err').read()[-2500:])
PY
done
