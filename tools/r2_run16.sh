#!/bin/bash
mkdir -p gpurun_out
for TI in 128 256 512; do
STTS_TAIL_TI=$TI timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2y_bench_ti$TI.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2y_bench_ti$TI.json")); print("TI=$TI", d["ms_per_step"], d["stage_ms_last_step"]["dec"])
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ms_tail|istft|synth_fir|pcm_kernel|refpad|split_planes|mrf_combine" -c 30 --csv --log-file gpurun_out/r2y_tail.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2y_tail.csv")) if len(r)>10 and r[0].isdigit()]
print([(r[4].split('(')[0].replace('stts::','')[:24], int(float(r[-1])/1000)) for r in rows[:30]])
PY
for PC in 1 3; do
STTS_PC_FUSED=$PC STTS_PC_TRACE=1 STTS_B200_LIB=tools/_build/libstts_b200_trace.so timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep PCTRACE | tee gpurun_out/r2y_pctrace_pc$PC.txt
done
