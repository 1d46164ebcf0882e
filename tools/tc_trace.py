"""Trace one CTA of a WN-in-layer-shaped conv (STTS_TC_TRACE=1) — GPU box tool."""
import os, sys
import numpy as np
os.environ["STTS_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from summertts_b200 import engine
rng = np.random.default_rng(0)
shape = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (384, 192, 5, 1))]
o, c, k, d = shape
w = (rng.standard_normal((o, k, c)) * 0.1).astype(np.float32)
rec = np.concatenate([np.array([o, c, k, d * (k - 1) // 2, d, 1], np.float32), w.ravel(), np.zeros(o, np.float32)]).astype(np.float32)
B, L = 64, (640 if len(sys.argv) < 6 else int(sys.argv[5]))
seg = (np.arange(B + 1) * L).astype(np.int32)
x = rng.standard_normal((B * L, c)).astype(np.float32)
for _ in range(2):
    engine.test_conv1d(rec, x, use_tc=1, seg_off=seg, epi=(4 if o == 2 * c and k == 5 else 0))
