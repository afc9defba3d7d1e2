#!/bin/bash
mkdir -p gpurun_out
for wl in ${WORKLOADS:-cifar6 mlp8192 conv1d_seq}; do
  for prec in ${PRECS:-tf32}; do
    timeout 600 python bench.py --workload $wl --precision $prec --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline ${EXTRA} > gpurun_out/bench_${wl}_${prec}.log 2> gpurun_out/bench_${wl}_${prec}.err; echo "$wl $prec exit $?"
    tail -2 gpurun_out/bench_${wl}_${prec}.err | cut -c1-300
    python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_${wl}_${prec}.log").read().strip().splitlines()[-1])
    print({k:l[k] for k in ("value","ms_per_step","gpu_launches","step_tflops","final_loss")}, "e2e", l["e2e"] and l["e2e"]["value"])
    for r in l["roofline_all"]: print("  ", r["kernel"], r["us"], r["frac"], r["unit"])
except Exception as e: print("parse fail", e)
PY
  done
done
