#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rbpair.py tests/test_dropin.py -m gpu -q -x --tb=short 2>&1 | tail -25 > gpurun_out/r2o_pytest.txt
cat gpurun_out/r2o_pytest.txt
if grep -q "failed\|rror" gpurun_out/r2o_pytest.txt; then
  STTS_PC_FUSED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line 2>&1 | tail -5
  exit 0
fi
for T in 1 2; do
STTS_TC_VERBOSE=4 timeout 300 python bench.py --no-cpu-baseline --steps 5 --tensor $T > gpurun_out/r2o_bench_t$T.json 2> gpurun_out/r2o_bench_t$T.err
grep pc_conv gpurun_out/r2o_bench_t$T.err | head -4
python - <<PY
import json
d=json.load(open("gpurun_out/r2o_bench_t$T.json")); print("tensor$T", d["ms_per_step"], {k:(round(v["ms"],2), round(v["tflops"])) for k,v in d["conv_classes"].items()}, d["stage_ms_last_step"])
PY
done
