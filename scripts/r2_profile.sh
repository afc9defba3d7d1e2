#!/bin/bash
# ncu captures for profiles/: one cifar6 step (all tcgen05 conv kernels), the 8192^3 bf16 dense GEMM, the small-C conv3d kernels.
# Reports stay in /tmp on the GPU box (they exceed gpurun's 64 MiB copy-back limit); the summaries are produced there.
mkdir -p gpurun_out /tmp/prof
CMD="python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_cifar6_b1024.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
SKIP=$(python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/r2_launches_cifar6_b1024.csv') if not l.startswith('==')]
r=list(csv.DictReader(lines))
idx=[i for i,x in enumerate(r) if 'accum_update' in x['Kernel Name']]
print(idx[-2]+1, idx[-1]-idx[-2])
PY
)
set -- $SKIP
echo "skip $1 count $2"
ncu --set full --clock-control none --import-source on -s $1 -c $2 -f -o /tmp/prof/cifar6_step $CMD > gpurun_out/ncu_full1.log 2>&1; echo "full cifar6 exit $?"
python scripts/ncu_summary.py /tmp/prof/cifar6_step.ncu-rep gpurun_out/r2_ncu_cifar6_kernels.json cifar6 "$CMD" > gpurun_out/r2_ncu_cifar6_kernels.txt 2>&1
rm -f gpurun_out/r2_ncu_cifar6_stalls.txt
for k in "conv_tc_kernel<\(int\)64" "conv_wgrad_gs_kernel" "conv_wgrad_tc_kernel<\(int\)128" "conv_first_fprop" "conv_first_wgrad_kernel" "head_big_kernel"; do
  python scripts/ncu_top.py /tmp/prof/cifar6_step.ncu-rep "$k" 0 18 >> gpurun_out/r2_ncu_cifar6_stalls.txt 2>&1
done
CMD2="python scripts/gemm_bench.py 8192 8192 8192 1 0 1"
ncu --set full --clock-control none -k regex:tc_gemm_kernel -s 3 -c 1 -f -o /tmp/prof/gemm8192 $CMD2 > gpurun_out/ncu_full2.log 2>&1; echo "full gemm exit $?"
python scripts/ncu_summary.py /tmp/prof/gemm8192.ncu-rep gpurun_out/r2_ncu_gemm8192_bf16.json plain "$CMD2" > gpurun_out/r2_ncu_gemm8192_bf16.txt 2>&1
CMD3="python bench.py --workload conv3d_video --precision bf16 --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --no-roofline"
ncu --set full --clock-control none --import-source on -k regex:"conv_first|mse_nxc" -s 9 -c 3 -f -o /tmp/prof/conv3d_bf16 $CMD3 > gpurun_out/ncu_full3.log 2>&1; echo "full conv3d exit $?"
python scripts/ncu_summary.py /tmp/prof/conv3d_bf16.ncu-rep gpurun_out/r2_ncu_conv3d_bf16.json plain "$CMD3" > gpurun_out/r2_ncu_conv3d_bf16.txt 2>&1
du -sh gpurun_out
