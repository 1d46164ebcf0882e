#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "shipped or throughput" 2>&1 | tail -60 > gpurun_out/r2n_pytest.txt; cat gpurun_out/r2n_pytest.txt
timeout 400 python tools/lsb_stats.py > gpurun_out/r2n_accuracy.txt 2>&1; cat gpurun_out/r2n_accuracy.txt
