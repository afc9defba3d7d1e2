#!/bin/bash
# round-2 baseline probe: cifar6 at the full batch (1024) + launch list + ncu --set full of the tcgen05 conv kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
for prec in bf16 tf32; do
  timeout 300 python bench.py --workload cifar6 --batch 1024 --precision $prec --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline > gpurun_out/r2_cifar6_b1024_$prec.log 2> gpurun_out/r2_cifar6_b1024_$prec.err; echo "$prec exit $?"
  cat gpurun_out/r2_cifar6_b1024_$prec.log | cut -c1-400
done
timeout 120 python scripts/conv_timeline.py bf16 > gpurun_out/r2_conv_timeline_bf16.log 2>&1; tail -30 gpurun_out/r2_conv_timeline_bf16.log
CMD="python bench.py --workload cifar6 --batch 1024 --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cifar6_b1024.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 40 -c 14 -f -o gpurun_out/r2_prof_cifar6_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
