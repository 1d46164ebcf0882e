#!/bin/bash
# tools/r2_final2.sh — second (last) GPU call of round 2: the two g2p kernels (streaming / cluster-resident) through the parity tests
# and the throughput tool, the whole GPU suite and a short bench on the final library, ncu of the two g2p kernels, sanitizer.
T=r2g
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/${T}_timeline.txt; }
stamp start
timeout 200 python -m pytest tests/test_g2p.py -m gpu -q -rs > gpurun_out/${T}_pytest_g2p.txt 2>&1; tail -15 gpurun_out/${T}_pytest_g2p.txt
stamp pytest-g2p
timeout 120 python tools/g2p_bench.py > gpurun_out/${T}_g2p_bench.json 2> gpurun_out/${T}_g2p_bench.err; cat gpurun_out/${T}_g2p_bench.json; tail -3 gpurun_out/${T}_g2p_bench.err
stamp g2p-bench
timeout 400 python -m pytest tests -m gpu -q -rs --deselect tests/test_g2p.py > gpurun_out/${T}_pytest_rest.txt 2>&1; tail -6 gpurun_out/${T}_pytest_rest.txt
stamp pytest-rest
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -3 gpurun_out/${T}_smoke.txt
stamp smoke
timeout 200 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_short.json 2> gpurun_out/${T}_bench_short.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_short.json")); print("bench", "%.3f ms  %.1f Msamples/s  e2e %.1f"%(d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), d["roofline"]["kernel_class"], "%.3f"%d["roofline"]["frac"])
except Exception as e: print("bench ERR", e)
PY
stamp bench
for k in 0 1; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:g2p_ -s 2 -c 1 -f -o gpurun_out/${T}_ncu_full_g2p_k$k python tools/g2p_probe.py $k 592 > gpurun_out/${T}_ncu_g2p_k$k.log 2>&1
  R=gpurun_out/${T}_ncu_full_g2p_k$k.ncu-rep
  [ -f $R ] && python tools/ncu_summary.py $R gpurun_out/${T}_ncu_full_g2p_k${k}_summary.csv > gpurun_out/${T}_ncu_key_metrics_g2p_k$k.txt 2>&1
  tail -2 gpurun_out/${T}_ncu_g2p_k$k.log
done
stamp ncu-g2p
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:relattn|ms_tail" -s 20 -c 2 -f -o gpurun_out/${T}_ncu_full_misc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_misc.log 2>&1
[ -f gpurun_out/${T}_ncu_full_misc.ncu-rep ] && python tools/ncu_summary.py gpurun_out/${T}_ncu_full_misc.ncu-rep gpurun_out/${T}_ncu_full_misc_summary.csv > gpurun_out/${T}_ncu_key_metrics_misc.txt 2>&1
stamp ncu-misc
for k in 0 1; do
  timeout 120 compute-sanitizer --tool racecheck python tools/g2p_probe.py $k 40 > gpurun_out/${T}_sanitizer_racecheck_g2p_k$k.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_racecheck_g2p_k$k.txt
  timeout 120 compute-sanitizer --tool memcheck python tools/g2p_probe.py $k 40 > gpurun_out/${T}_sanitizer_memcheck_g2p_k$k.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_memcheck_g2p_k$k.txt
done
stamp sanitizer
ls -la gpurun_out | grep ${T}_
