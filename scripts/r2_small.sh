#!/bin/bash
mkdir -p gpurun_out
python bench.py --batch 128 --steps 50 --no-cpu-baseline --no-configs --no-roofline > gpurun_out/r2_bench_b128.json 2> gpurun_out/r2_bench_b128.err; echo "exit $?"
python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_b128.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','step_tflops','gpu_launches','repeats')}, 'e2e', l['e2e'] and l['e2e']['value'])
"
CMD="python bench.py --batch 128 --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_cifar6_b128.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
TL_BATCH=128 python scripts/step_timeline.py cifar6 bf16 > gpurun_out/r2_timeline_b128.log 2>&1; tail -80 gpurun_out/r2_timeline_b128.log
