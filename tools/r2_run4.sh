#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2e_pytest.txt
cat gpurun_out/r2e_pytest.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rb_pair -c 72 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2e_launches.csv")) if len(r)>10 and r[0].isdigit()]
print("rb_pair launches", len(rows))
for r in rows[:18]: print(r[4][:40], r[-1], r[-2], r[-3])
PY
for K in 11 3; do
  for T in 1 2; do
    STTS_B200_LIB=tools/_build/libstts_b200_trace.so STTS_RB_TRACE=$K timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tensor $T 2>&1 >/dev/null | grep -A4 RBTRACE > gpurun_out/r2e_trace_k${K}_t${T}.txt
    cat gpurun_out/r2e_trace_k${K}_t${T}.txt
  done
done
