#!/bin/bash
# tools/final_run.sh — the round-end GPU recipe (run on the B200 box through gpurun, from the repo root):
#   parity tests, smoke, accuracy table, ncu full captures of the dominant kernels -> profiles/traffic.json,
#   the bench line, and the ncu launch list of the same bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest.txt
cat gpurun_out/final_pytest.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
tail -2 gpurun_out/final_smoke.txt
timeout 150 python tools/lsb_stats.py > gpurun_out/final_lsb.txt 2>&1
cat gpurun_out/final_lsb.txt
# conv_tc launch index inside the first step: 28 = WN in-layer (k5, 192->384), 29 = res_skip; 75/76 = ResBlock1 conv1/conv2 k7 at
# 64 channels; 94/95 = the same at 32 channels
for spec in "flow 28" "dec1 75" "dec2 94"; do
  set -- $spec
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $2 -c 2 -f \
      -o gpurun_out/prof_r1_final_$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
done
python tools/ncu_traffic.py profiles/traffic.json wn_in,wn_rs=gpurun_out/prof_r1_final_flow.ncu-rep \
    dec_rb=gpurun_out/prof_r1_final_dec1.ncu-rep dec_rb=gpurun_out/prof_r1_final_dec2.ncu-rep > gpurun_out/final_traffic.txt 2>&1
cp profiles/traffic.json gpurun_out/traffic.json
cat gpurun_out/final_traffic.txt
timeout 400 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
cat gpurun_out/bench_r1_final.json
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out | tail -20
