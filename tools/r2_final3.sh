#!/bin/bash
# tools/r2_final3.sh — closing check of the final library (g2p kernel chosen per call): g2p parity tests, g2p throughput of the three
# settings, the rest of the GPU suite, smoke().
T=r2h
mkdir -p gpurun_out
t0=$(date +%s); stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/${T}_timeline.txt; }
stamp start
timeout 200 python -m pytest tests/test_g2p.py -m gpu -q -rs > gpurun_out/${T}_pytest_g2p.txt 2>&1; tail -15 gpurun_out/${T}_pytest_g2p.txt
stamp pytest-g2p
timeout 120 python tools/g2p_bench.py > gpurun_out/${T}_g2p_bench.json 2> gpurun_out/${T}_g2p_bench.err; cat gpurun_out/${T}_g2p_bench.json; tail -3 gpurun_out/${T}_g2p_bench.err
stamp g2p-bench
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -3 gpurun_out/${T}_smoke.txt
stamp smoke
timeout 400 python -m pytest tests -m gpu -q -rs > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -6 gpurun_out/${T}_pytest_gpu.txt
stamp pytest-all
