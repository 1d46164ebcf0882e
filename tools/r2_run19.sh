#!/bin/bash
mkdir -p gpurun_out
for D in 0 1 2 3 4 7; do
  STTS_PC_FUSED=3 STTS_PC_DBG=$D timeout -s KILL 120 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2z_dbg$D.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2z_dbg$D.json")); print("dbg $D", round(d["ms_per_step"],3), "wn_in", d["conv_classes"]["wn_in"], "fallbacks", d["tensor_fallbacks"])
except Exception as e: print("dbg $D failed", e)
PY
done
timeout -s KILL 150 python bench.py --no-cpu-baseline --steps 3 --tensor 2 > gpurun_out/r2z_t2.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2z_t2.json")); print("tensor2", round(d["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["conv_classes"].items()}, "fallbacks", d["tensor_fallbacks"])
PY
