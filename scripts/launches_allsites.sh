#!/bin/bash
# GPU box: ncu per-launch durations of one scoreAllSites call at 8 M reads
mkdir -p gpurun_out
CMD="python scripts/bench_allsites.py 8000000"
$CMD > gpurun_out/plain_as.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_as.csv $CMD > gpurun_out/ncu_launch_as.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_as.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[h+1:]:
    if len(r)>vi:
        v=float(r[vi].replace(',',''))
        if r[ui]=='ns': v/=1000.0
        elif r[ui]=='ms': v*=1000.0
        seq.append((r[ki][:60], v))
idx=[i for i,(b,c) in enumerate(seq) if 'k_batch_stats' in b]
for b,c in seq[idx[-1]:]: print('%-62s %9.1f us'%(b,c))
PY
