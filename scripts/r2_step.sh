#!/bin/bash
# probe of the new conv kernels + the subset of GPU tests that exercises them + a short bench + launch list
mkdir -p gpurun_out
if [ "${PROBE:-first misc parity}" != "none" ]; then
timeout 600 python scripts/conv2_probe.py ${PROBE:-first misc parity} > gpurun_out/conv2_probe.log 2>&1; echo "probe exit $?"; grep -c "^ok" gpurun_out/conv2_probe.log; grep -v "^ok" gpurun_out/conv2_probe.log | tail -40
fi
timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/r2_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2_smoke.log | cut -c1-400
for wl in ${WORKLOADS:-cifar6}; do
timeout 600 python bench.py --workload $wl --steps 10 --no-configs --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "bench $wl exit $?"; tail -5 gpurun_out/r2_bench_$wl.err
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_$wl.json').read().strip().splitlines()[-1])
    print({k:l.get(k) for k in ('value','ms_per_step','step_tflops','gpu_launches','final_loss','repeats','dtype')}, 'e2e', l['e2e'] and l['e2e']['value'])
    for r in l.get('roofline_all') or []: print('  ', r['kernel'], r['us'], r['frac'], r['bound'], r.get('tflops'))
except Exception as e: print('parse fail', e)
PY
done
if [ -n "${NCU_LIST}" ]; then
CMD="python bench.py --workload ${NCU_LIST} --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --no-roofline ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_${NCU_LIST}.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
fi
