#!/bin/bash
mkdir -p gpurun_out
P="timeout -s KILL 75 python tools/pc_probe.py single_speaker_fast"
STTS_PC_FUSED=0 STTS_RB_FUSED=0 $P gpurun_out/probe_ref.npy 2>&1 | tail -2
[ -f gpurun_out/probe_ref.npy ] || { echo "reference probe failed"; exit 1; }
export PROBE_REF=gpurun_out/probe_ref.npy
$P 2>&1 | tail -2; rc2=${PIPESTATUS[0]}
PROBE_TENSOR=2 $P 2>&1 | tail -2; rc4=${PIPESTATUS[0]}
echo "rc: $rc2 $rc4"
if [ "$rc2" != "0" ] || [ "$rc4" != "0" ]; then echo "probe failed: stop"; exit 1; fi
echo "== rbpair"; timeout -s KILL 300 python -m pytest tests/test_gpu_rbpair.py -x -q -m gpu 2>&1 | tail -2
for CFG in "default:" "t2:BENCH_TENSOR=2"; do
  NAME=${CFG%%:*}; ENVS=${CFG#*:}
  T=""; case "$ENVS" in *BENCH_TENSOR=2*) T="--tensor 2";; esac
  env $ENVS timeout -s KILL 150 python bench.py --no-cpu-baseline --steps 5 $T > gpurun_out/r2u_bench_$NAME.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2u_bench_$NAME.json")); print("$NAME", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms_last_step"].items()}, {k:round(v["ms"],3) for k,v in d["conv_classes"].items()})
except Exception as e: print("$NAME failed", e)
PY
done
