#!/bin/bash
# BASELINE configs[3] (Conv1D sequence net, Conv3D video block): step time per precision + ncu launch lists
mkdir -p gpurun_out
for wl in conv1d_seq conv3d_video; do
  for prec in tf32 bf16; do
    timeout 300 python bench.py --workload $wl --precision $prec --steps 10 --no-configs --no-cpu-baseline --no-roofline > gpurun_out/cfg3_${wl}_${prec}.json 2> gpurun_out/cfg3_${wl}_${prec}.err
    echo "$wl $prec exit $?"; python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/cfg3_${wl}_${prec}.json').read().strip().splitlines()[-1])
    print({k:l.get(k) for k in ('value','ms_per_step','step_tflops','gpu_launches')}, 'e2e', l['e2e'] and l['e2e']['value'])
except Exception as e: print('parse fail', e)
PY
    CMD="python bench.py --workload $wl --precision $prec --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --no-roofline"
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/cfg3_launches_${wl}_${prec}.csv $CMD > /dev/null 2>&1
    echo "ncu exit $?"
  done
done
