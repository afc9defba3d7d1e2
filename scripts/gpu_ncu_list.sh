#!/bin/bash
# launch list of one bench run (per-launch device time, cold-cache & serialised: compare shares)
mkdir -p gpurun_out
CMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none ${NCU_EXTRA} -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; rows=rows[i+1:]; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
seq=[]
for r in rows:
    try: v=float(r[vi].replace(',',''))
    except: continue
    seq.append((r[ki][:70],v))
# last 31*? take the last step: find graph portion = last N launches
tail=seq[-400:]
tot=collections.Counter(); cnt=collections.Counter()
for k,v in tail: tot[k]+=v; cnt[k]+=1
s=sum(tot.values())
for k,v in tot.most_common(25): print(f"{v/1000:9.1f} us {100*v/s:5.1f}%  n={cnt[k]:4d}  {k}")
PY
