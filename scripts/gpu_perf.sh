#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "exit $?" >> gpurun_out/gemm_bench.log
cat gpurun_out/gemm_bench.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tf32.log 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_tf32.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','step_tflops')}, l['e2e']['value'])
for r in l['roofline_all']: print(r['kernel'], r['us'], r['frac'], r['unit'])
PY
