#!/bin/bash
mkdir -p gpurun_out
echo "== parity default"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rbpair.py tests/test_gpu_conv_ops.py -x -q -m gpu 2>&1 | tail -3
echo "== parity PC_FUSED=3"; STTS_PC_FUSED=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for CFG in "default:" "pc3:STTS_PC_FUSED=3" "tma:STTS_TILE_TMA=1" "pc3tma:STTS_PC_FUSED=3 STTS_TILE_TMA=1"; do
  NAME=${CFG%%:*}; ENVS=${CFG#*:}
  env $ENVS timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2z_bench_$NAME.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench_$NAME.json")); print("$NAME", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms_last_step"].items()}, {k:round(v["ms"],3) for k,v in d["conv_classes"].items()} if isinstance(d["conv_classes"],dict) else "")
PY
done
for PC in 1 3; do
STTS_PC_FUSED=$PC STTS_PC_TRACE=1 STTS_B200_LIB=tools/_build/libstts_b200_trace.so timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep PCTRACE | tee gpurun_out/r2z_pctrace_pc$PC.txt
done
