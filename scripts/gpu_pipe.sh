#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for pf in 0 1; do for head in 0 1; do
  OKL_PREFETCH=$pf OKL_FUSE_HEAD=$head timeout 300 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/b_${pf}_${head}.log 2> gpurun_out/b_${pf}_${head}.err
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/b_${pf}_${head}.log').read().strip().splitlines()[-1])
    print('prefetch=$pf head=$head', {k:l[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', l['e2e'] and (l['e2e']['value'], l['e2e']['ms_per_step']))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/b_${pf}_${head}.err').read()[-1500:])
PY
done; done; done
