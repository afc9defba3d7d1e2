#!/bin/bash
# exactly what the driver runs at round end: default bench line + reference arm, with wall-clock per command
mkdir -p gpurun_out
s=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_default_bench.json 2> gpurun_out/r2_default_bench.err; echo "default bench exit $? in $(( $(date +%s) - s )) s"
tail -3 gpurun_out/r2_default_bench.err
s=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/r2_default_ref.json 2> gpurun_out/r2_default_ref.err; echo "reference arm exit $? in $(( $(date +%s) - s )) s"
tail -3 gpurun_out/r2_default_ref.err
cut -c1-600 gpurun_out/r2_default_ref.json
