#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 700 --csv --log-file gpurun_out/r2s_launches_multi.csv python bench.py --model multi_speakers --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2s_launches_multi.csv")) if len(r)>10 and r[0].isdigit()]
t=collections.defaultdict(float); n=collections.Counter()
for r in rows:
    name=r[4].split('(')[0][:50]+" grid="+r[8] if False else r[4].split('(')[0][:50]
    t[name]+=float(r[-1])/1000; n[name]+=1
tot=sum(t.values()); print("launches",len(rows),"total us",round(tot))
for k,v in sorted(t.items(), key=lambda x:-x[1])[:14]: print("%-52s %4d %9.1f us %5.1f%%"%(k,n[k],v,100*v/tot))
# sequence of big launches
print([ (r[4].split('(')[0].replace('void stts::','')[:22], int(float(r[-1])/1000)) for r in rows if float(r[-1])>150000][:80])
PY
