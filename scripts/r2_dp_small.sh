#!/bin/bash
# secondary single-layer configs under data parallel (their whole gradient is one small bucket)
N=${N:-2}
mkdir -p gpurun_out
for wl in ${WLS:-conv1d_seq conv3d_video conv2d_single mnist_cnn}; do
  prec=bf16; [ $wl = conv2d_single ] && prec=tf32; [ $wl = mnist_cnn ] && prec=tf32
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload $wl --precision $prec --steps 50 --no-configs --no-cpu-baseline --no-e2e --no-roofline > gpurun_out/dps_$wl.json 2> gpurun_out/dps_$wl.err
  echo "$wl exit $?"; python -c "
import json
l=json.loads(open('gpurun_out/dps_$wl.json').read().strip().splitlines()[-1]); print('$wl', l['value'], l['ms_per_step'])"
done
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests -m gpu -x -q -k "dp" 2>&1 | tail -2
