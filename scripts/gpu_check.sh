#!/bin/bash
# One gpurun call: tests, smoke and a short bench, each under its own timeout, logs in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','final_loss')}, 'e2e', l['e2e'] and l['e2e']['value'], 'cpu', l['cpu_baseline'] and l['cpu_baseline']['value'])
for r in l['roofline_all']: print(r['kernel'], r['us'], r['frac'], r['unit'])
PY
