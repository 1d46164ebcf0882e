#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/r2p_pytest.txt
cat gpurun_out/r2p_pytest.txt
if grep -q "failed\|rror" gpurun_out/r2p_pytest.txt; then exit 0; fi
for T in 1 2; do
STTS_TC_VERBOSE=4 timeout 300 python bench.py --no-cpu-baseline --steps 5 --tensor $T > gpurun_out/r2p_bench_t$T.json 2> gpurun_out/r2p_bench_t$T.err
grep "pc_conv: cluster" gpurun_out/r2p_bench_t$T.err | head -2
python - <<PY
import json
d=json.load(open("gpurun_out/r2p_bench_t$T.json")); print("tensor$T", d["ms_per_step"], {k:(round(v["ms"],2), round(v["tflops"])) for k,v in d["conv_classes"].items()})
PY
done
STTS_PC_CLUSTER=1 timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2p_bench_noclu.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2p_bench_noclu.json")); print("nocluster", d["ms_per_step"], {k:(round(v["ms"],2), round(v["tflops"])) for k,v in d["conv_classes"].items()})
PY
