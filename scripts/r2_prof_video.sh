#!/bin/bash
# ncu --set full of one conv3d_video BF16 step (small-K tcgen05 fprop / wgrad, bf16 MSE) + source-level stall summaries
mkdir -p gpurun_out /tmp/prof
CMD3="python bench.py --workload conv3d_video --precision bf16 --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --no-roofline"
$CMD3 > gpurun_out/plain3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"conv_first|mse_nxc" -s 9 -c 3 -f -o /tmp/prof/conv3d_bf16 $CMD3 > gpurun_out/ncu_full3.log 2>&1; echo "full conv3d exit $?"
python scripts/ncu_summary.py /tmp/prof/conv3d_bf16.ncu-rep gpurun_out/r2_ncu_conv3d_bf16.json plain "$CMD3" > gpurun_out/r2_ncu_conv3d_bf16.txt 2>&1
cat gpurun_out/r2_ncu_conv3d_bf16.txt
for k in "conv_first_fprop" "conv_first_wgrad_kernel"; do
  python scripts/ncu_top.py /tmp/prof/conv3d_bf16.ncu-rep "$k" 0 22 >> gpurun_out/r2_ncu_conv3d_bf16_stalls.txt 2>&1
done
