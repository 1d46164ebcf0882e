#!/bin/bash
# tools/r2_final.sh — the round-2 closing GPU run (one gpurun call, shipped weights present), most important first so that a
# clamped call still leaves the parity result and the bench line: parity suite -> bench line -> ncu launch list -> ncu full
# captures of the fused ResBlock1-pair and WaveNet kernels -> bench lines of the other modes / models -> g2p throughput ->
# accuracy table of the three arithmetic modes -> sanitizer.  Everything lands in gpurun_out/r2f_*.
T=r2f
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/${T}_timeline.txt; }
stamp start; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader | tee -a gpurun_out/${T}_timeline.txt
timeout 600 python -m pytest tests -m gpu -q -rs --durations=8 > gpurun_out/${T}_pytest_full.txt 2>&1; tail -25 gpurun_out/${T}_pytest_full.txt
stamp pytest
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json
stamp bench
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1
stamp launch-list
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair -s 21 -c 10 -f -o gpurun_out/${T}_ncu_full_rbpair python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_rb.log 2>&1
stamp ncu-rbpair
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pc_kernel -s 34 -c 4 -f -o gpurun_out/${T}_ncu_full_pc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_pc.log 2>&1
stamp ncu-pc
for k in rbpair pc; do
  R=gpurun_out/${T}_ncu_full_$k.ncu-rep
  [ -f $R ] || continue
  python tools/ncu_summary.py $R gpurun_out/${T}_ncu_full_${k}_summary.csv > gpurun_out/${T}_ncu_key_metrics_$k.txt 2>&1
  ncu -i $R --page raw --csv > gpurun_out/${T}_ncu_full_${k}_raw.csv 2>/dev/null
  [ $(stat -c %s $R) -gt 24000000 ] && rm -f $R
done
stamp ncu-summaries
timeout 200 python bench.py --tensor 2 --no-cpu-baseline > gpurun_out/${T}_bench_t2.json 2>/dev/null
timeout 200 python bench.py --model single_speaker_mid --no-cpu-baseline > gpurun_out/${T}_bench_mid.json 2>/dev/null
timeout 200 python bench.py --model multi_speakers --no-cpu-baseline > gpurun_out/${T}_bench_multi.json 2>/dev/null
timeout 200 python bench.py --model multi_speakers --durations model --no-cpu-baseline > gpurun_out/${T}_bench_multi_modeldur.json 2>/dev/null
python - <<PY
import json
for f in ("bench","bench_t2","bench_mid","bench_multi","bench_multi_modeldur"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%f)); print(f, "%.3f ms  %.1f Msamples/s  e2e %.1f"%(d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), d["roofline"]["kernel_class"], "%.3f"%d["roofline"]["frac"], {k:round(v["ms"],2) for k,v in d["conv_classes"].items()})
    except Exception as e: print(f, "ERR", e)
PY
stamp benches
timeout 120 python tools/g2p_bench.py > gpurun_out/${T}_g2p_bench.json 2> gpurun_out/${T}_g2p_bench.err; cat gpurun_out/${T}_g2p_bench.json
stamp g2p
timeout 300 python tools/lsb_stats.py > gpurun_out/${T}_accuracy.txt 2>&1; cat gpurun_out/${T}_accuracy.txt
stamp accuracy
timeout 200 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_memcheck_smoke.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_memcheck_smoke.txt
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_g2p.py -m gpu -q -k "synthetic or ragged" > gpurun_out/${T}_sanitizer_racecheck_g2p.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_racecheck_g2p.txt
timeout 200 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_racecheck_smoke.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_racecheck_smoke.txt
stamp sanitizer
ls -la gpurun_out | tail -40
