#!/bin/bash
mkdir -p gpurun_out
for bn in 0 32 64 128; do for sp in 0 8 16 32; do
  OKL_TC_BN=$bn OKL_TC_SPLITS=$sp timeout 120 python bench.py --steps 600 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/sw.log 2>gpurun_out/sw.err
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/sw.log').read().strip().splitlines()[-1])
    print('bn=$bn splits=$sp', l['ms_per_step'], [ (r['kernel'][:13], r['us']) for r in l['roofline_all'][2:]])
except Exception as e:
    print('bn=$bn splits=$sp ERR', open('gpurun_out/sw.err').read()[-300:])
PY
done; done
