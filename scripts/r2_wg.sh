#!/bin/bash
# weight-gradient probe: parity + timing + kernel-only times (ncu launch list of the perf part)
mkdir -p gpurun_out
timeout 600 python scripts/conv2_probe.py wgrad > gpurun_out/wg_probe.log 2>&1; echo "probe exit $?"; grep -v "^ok" gpurun_out/wg_probe.log | tail -40; grep -c "^ok" gpurun_out/wg_probe.log
PROBE_SKIP_PARITY=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad|chansum|tc_gemm|transpose" -c 400 --csv --log-file gpurun_out/wg_launches.csv python scripts/conv2_probe.py wgrad > gpurun_out/wg_ncu.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv
rows=list(csv.DictReader([l for l in open('gpurun_out/wg_launches.csv') if l.startswith('"')]))
last=None
for r in rows:
    n=r['Kernel Name'][:60]; g=r['Grid Size']; key=(n,g)
    if key!=last: print(f"{float(r['Metric Value'].replace(',',''))/1000:8.1f} us  grid {g:>14s} {n}")
    last=key
PY
