#!/bin/bash
mkdir -p gpurun_out
for D in 0 7; do
STTS_PC_DBG=$D STTS_PC_FUSED=3 STTS_PC_TRACE=1 STTS_PC_TRACE_FILE=gpurun_out/pc_tl_d$D STTS_B200_LIB=tools/_build/libstts_b200_trace.so timeout -s KILL 120 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep PCTRACE | head -3
done
ls gpurun_out/pc_tl*
