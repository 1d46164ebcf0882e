#!/bin/bash
# round-2 dev run: fused ResBlock1-pair kernel — op tests, full GPU suite, bench (random-init weights when the shipped ones are not pushed)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rbpair.py -x -q 2>&1 | tail -25 > gpurun_out/r2c_rbpair.txt
cat gpurun_out/r2c_rbpair.txt
if grep -q "passed" gpurun_out/r2c_rbpair.txt && ! grep -q "failed" gpurun_out/r2c_rbpair.txt; then
  timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c_pytest.txt
  cat gpurun_out/r2c_pytest.txt
  STTS_TC_VERBOSE=60 timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
  cat gpurun_out/r2c_bench.json
  grep rb_pair gpurun_out/r2c_bench.err | head -20
  timeout 300 python bench.py --no-cpu-baseline --steps 5 --tensor 2 > gpurun_out/r2c_bench_t2.json 2> gpurun_out/r2c_bench_t2.err
  cat gpurun_out/r2c_bench_t2.json
  STTS_RB_FUSED=0 timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2c_bench_unfused.json 2> /dev/null
  python - <<'PY'
import json
for f in ("r2c_bench","r2c_bench_t2","r2c_bench_unfused"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], {k:(round(v["ms"],3),round(v["tflops"],1)) for k,v in d["conv_classes"].items()})
    except Exception as e: print(f, "ERR", e)
PY
fi
