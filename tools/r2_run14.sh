#!/bin/bash
# 2-GPU sanity: weak and strong scaling lines + the reference arm under torchrun (rank 0 only works)
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_n2_weak.json 2> gpurun_out/r2w_n2_weak.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --scaling strong --model multi_speakers > gpurun_out/r2w_bench_n2_strong_multi.json 2> gpurun_out/r2w_n2_strong.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scaling strong --model multi_speakers > gpurun_out/r2w_bench_n1_strong_multi.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2w_bench_n2_weak","r2w_bench_n2_strong_multi","r2w_bench_n1_strong_multi"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["scaling"], "%.3f ms"%d["ms_per_step"], "%.1f M/s"%(d["value"]/1e6), "e2e %.1f"%(d["e2e"]["value"]/1e6), d["config"]["batch_per_gpu"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.json"%f).read()[-500:])
PY
tail -3 gpurun_out/r2w_n2_weak.err
