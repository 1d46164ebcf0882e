#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -25 > gpurun_out/r2q_pytest.txt
cat gpurun_out/r2q_pytest.txt
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests")
from parity_util import TEST_TXT_IDS
from summertts_b200 import binfmt, engine
E = engine.SynthesizerTrn(binfmt.synthetic_model(seed=11))
ids = (TEST_TXT_IDS[:-1] * 8) + [1]
for _ in range(3): w = E.infer_ids(ids)
print("one-shot ms", E.last_timing()["total"], "samples", w.size)
for ch in (32, 64, 128, 256):
    for _ in range(2): c, f = E.infer_stream(ids, chunk_frames=ch)
    print("chunk", ch, "chunks", len(c), "first chunk ms %.3f" % f)
PY
timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2q_bench.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2q_bench.json")); print("bench", d["ms_per_step"], {k:(round(v["ms"],2), round(v["tflops"])) for k,v in d["conv_classes"].items()})
PY
