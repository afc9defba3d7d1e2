#!/bin/bash
# ncu --set full capture of selected kernels of a short bench run: KREGEX=... SKIP=... COUNT=...
mkdir -p gpurun_out
CMD="python bench.py --workload ${WL:-mnist_cnn} --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-tc_gemm} -s ${SKIP:-12} -c ${COUNT:-3} -f -o gpurun_out/${OUT:-prof} $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log
