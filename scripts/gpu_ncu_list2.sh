#!/bin/bash
mkdir -p gpurun_out
for wl in ${WORKLOADS}; do
CMD="python bench.py --workload $wl --precision ${PREC:-tf32} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_$wl.csv $CMD > gpurun_out/ncu.log 2>&1
echo "== $wl ncu exit $?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$wl.csv')) if len(r)>5]
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; rows=rows[i+1:]; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[]
for r in rows:
    try: seq.append((r[ki][:86],float(r[vi].replace(',',''))))
    except: pass
tail=seq[len(seq)//2:]
tot=collections.Counter(); cnt=collections.Counter()
for k,v in tail: tot[k]+=v; cnt[k]+=1
s=sum(tot.values())
for k,v in tot.most_common(16): print(f"{v/1000:10.1f} us {100*v/s:5.1f}%  n={cnt[k]:4d}  {k}")
PY
done
