#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rbpair.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2j_pytest.txt
cat gpurun_out/r2j_pytest.txt
if grep -q "failed\|error" gpurun_out/r2j_pytest.txt; then exit 0; fi
for T in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 5 --tensor $T > gpurun_out/r2j_bench_t$T.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2j_bench_t$T.json")); print("tensor$T", d["ms_per_step"], d["conv_classes"]["dec_rb"], d["stage_ms_last_step"])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rb_pair -c 36 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2j_launches.csv")) if len(r)>10 and r[0].isdigit()]
print("rb_pair launches", len(rows), [int(float(r[-1])/1000) for r in rows[:18]])
PY
