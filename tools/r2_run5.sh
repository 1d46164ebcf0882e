#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rbpair.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2g_pytest.txt
cat gpurun_out/r2g_pytest.txt
for D in 0 1; do
  for T in 1 2; do
    STTS_RB_DBG=$D STTS_B200_LIB=tools/_build/libstts_b200_trace.so STTS_RB_TRACE=11 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tensor $T 2>&1 >/dev/null | grep -A7 "RBTRACE C=32" > gpurun_out/r2g_trace_dbg${D}_t${T}.txt
    echo "== dbg $D tensor $T"; cut -c1-1500 gpurun_out/r2g_trace_dbg${D}_t${T}.txt
  done
done
for D in 0 1; do
STTS_RB_DBG=$D timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2g_bench_dbg$D.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_dbg$D.json")); print("dbg$D", d["ms_per_step"], d["conv_classes"]["dec_rb"])
PY
done
