#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_rbpair.py tests/test_gpu_parity.py tests/test_gpu_stream.py -m gpu -q -x --tb=short 2>&1 | tail -25 > gpurun_out/r2t_pytest.txt
cat gpurun_out/r2t_pytest.txt
if grep -q "failed\|rror" gpurun_out/r2t_pytest.txt; then exit 0; fi
timeout 300 python bench.py --model multi_speakers --no-cpu-baseline --steps 5 > gpurun_out/r2t_bench_multi.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2t_bench.json 2>/dev/null
python - <<PY
import json
for f in ("r2t_bench_multi","r2t_bench"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], {k:(round(v["ms"],2), round(v["tflops"])) for k,v in d["conv_classes"].items()})
PY
