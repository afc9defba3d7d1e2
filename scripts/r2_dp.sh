#!/bin/bash
# data-parallel run of the default bench (cifar6, global batch 1024) on N GPUs of one box
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 ${BENCH_ARGS} > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench n=$N exit $?"; tail -5 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
    print({k:l.get(k) for k in ('value','ms_per_step','step_tflops','gpu_launches','final_loss','repeats','scaling','dp_check')}, 'e2e', l['e2e'] and l['e2e']['value'])
    print('modes', l.get('modes'))
    for k,v in l['configs'].items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','step_tflops','precision','error')})
except Exception as e: print('parse fail', e)
PY
