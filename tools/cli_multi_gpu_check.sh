#!/bin/bash
# tools/cli_multi_gpu_check.sh — run on a box with >= 2 GPUs (gpurun --gpus 2): the demo CLI with --gpus 2 (one engine + one host
# thread per GPU inside ONE process: every engine sets the function attributes of the tensor-core kernels on its own device,
# stts_engine::device_setup) must write exactly the WAVs of --gpus 1; then the 2-GPU bench line (torchrun, weak scaling).
T=r2m
mkdir -p gpurun_out /tmp/cli2
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee gpurun_out/${T}_cli_2gpu.txt
MODEL=oracle/_ref/models/single_speaker_fast.bin
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests")
from parity_util import synth_ids
rng = np.random.default_rng(77)
lens = [128, 17, 64, 5, 96, 33, 128, 9, 77, 50, 21, 111]
open("/tmp/cli2/ids.txt", "w").write("".join(" ".join(map(str, synth_ids(rng, n))) + "\n" for n in lens))
PY
CLI=summertts_b200/bin/tts_b200_ids
timeout 120 $CLI --ids --gpus 1 /tmp/cli2/ids.txt $MODEL /tmp/cli2/a.wav >> gpurun_out/${T}_cli_2gpu.txt 2>&1; echo "rc gpus=1: $?" | tee -a gpurun_out/${T}_cli_2gpu.txt
timeout 120 $CLI --ids --gpus 2 /tmp/cli2/ids.txt $MODEL /tmp/cli2/b.wav >> gpurun_out/${T}_cli_2gpu.txt 2>&1; echo "rc gpus=2: $?" | tee -a gpurun_out/${T}_cli_2gpu.txt
same=0; n=0
for f in /tmp/cli2/a.wav_*.wav; do
  g=${f/a.wav_/b.wav_}; n=$((n+1))
  cmp -s $f $g && same=$((same+1))
done
echo "utterances: $n, bit-identical between --gpus 1 and --gpus 2: $same" | tee -a gpurun_out/${T}_cli_2gpu.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_2gpu.json")); print("2-GPU bench", d["n_gpus"], "%.3f ms  %.1f Msamples/s  e2e %.1f"%(d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), d["scaling"])
except Exception as e: print("2-GPU bench ERR", e)
PY
