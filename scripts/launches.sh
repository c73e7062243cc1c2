#!/bin/bash
# usage: bash scripts/launches.sh <tag>   -- GPU box: ncu per-launch durations of one bench step
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_$TAG.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[h+1:]:
    if len(r)>vi:
        v=float(r[vi].replace(',',''));
        if r[ui]=='ns': v/=1000.0
        elif r[ui]=='ms': v*=1000.0
        seq.append((r[ki][:50], v))
idx=[i for i,(b,c) in enumerate(seq) if 'k_build_index' in b or 'k_batch_stats' in b]
last=[i for i in idx if i>len(seq)//2][0] if idx else 0
tot=0
for b,c in seq[last:]:
    print('%-52s %9.1f us'%(b,c)); tot+=c
print('sum %.1f us'%tot)
PY
