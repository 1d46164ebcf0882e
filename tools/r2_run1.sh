#!/bin/bash
# round-2 run 1: baseline with shipped weights on the box + MMA micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 tools/_build/mma_bench2 > gpurun_out/r2a_mma_bench2.txt 2>&1
cat gpurun_out/r2a_mma_bench2.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2a_pytest.txt
cat gpurun_out/r2a_pytest.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cat gpurun_out/r2a_bench.json
nproc; python -c "import bench; print(bench.host_cores(), bench.cpu_model())"; cat /sys/fs/cgroup/cpu.max 2>/dev/null
