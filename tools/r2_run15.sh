#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -12 > gpurun_out/r2x_pytest.txt
cat gpurun_out/r2x_pytest.txt
if grep -q "failed\|rror" gpurun_out/r2x_pytest.txt; then exit 0; fi
timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2x_bench.json 2>/dev/null
STTS_TAIL_FUSED=0 timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2x_bench_notail.json 2>/dev/null
python - <<PY
import json
for f in ("r2x_bench","r2x_bench_notail"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], d["stage_ms_last_step"], d["gpu_launches"])
PY
