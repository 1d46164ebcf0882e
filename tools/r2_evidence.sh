#!/bin/bash
# round-2 evidence run (GPU box, shipped weights present): parity suite, accuracy tables of the three arithmetic modes, bench
# lines (default / throughput / mid / multi), ncu launch list + full captures of the fused ResBlock1-pair kernel, sanitizer.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
timeout 300 python tools/lsb_stats.py > gpurun_out/${TAG}_accuracy.txt 2>&1; cat gpurun_out/${TAG}_accuracy.txt
timeout 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --tensor 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_t2.json 2>/dev/null
timeout 300 python bench.py --model single_speaker_mid --no-cpu-baseline > gpurun_out/${TAG}_bench_mid.json 2>/dev/null
timeout 300 python bench.py --model single_speaker_mid --tensor 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_mid_t2.json 2>/dev/null
timeout 300 python bench.py --model multi_speakers --no-cpu-baseline > gpurun_out/${TAG}_bench_multi.json 2>/dev/null
timeout 300 python bench.py --model multi_speakers --durations model --no-cpu-baseline > gpurun_out/${TAG}_bench_multi_modeldur.json 2>/dev/null
python - <<PY
import json
for f in ("bench","bench_t2","bench_mid","bench_mid_t2","bench_multi","bench_multi_modeldur"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%f)); print(f, "%.3f ms  %.1f Msamples/s  e2e %.1f"%(d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), d["roofline"]["kernel_class"], "%.3f"%d["roofline"]["frac"], {k:round(v["ms"],2) for k,v in d["conv_classes"].items()})
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
for spec in "c64k7 21" "c32k7 30" "c32k11 33"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair -s $2 -c 1 -f -o gpurun_out/${TAG}_ncu_full_rbpair_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$1.log 2>&1
done
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.txt
ls -la gpurun_out | tail -30
