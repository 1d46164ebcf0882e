"""Shared helpers of the parity tests: model discovery, seeded inputs, error metrics."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_DIRS = [os.environ.get("STTS_MODEL_DIR", ""), os.path.join(ROOT, "oracle", "_ref", "models"),
              "/root/reference/models"]

# phoneme ids of /root/reference/test.txt through the reference's Chinese frontend
TEST_TXT_IDS = [int(x) for x in (
    "0 15 119 3 19 90 3 24 117 3 7 77 3 17 125 3 7 77 3 23 39 3 19 90 3 0 0 26 87 3 25 89 3 14 35 3 18 60 3 7 136 3 "
    "17 35 3 0 0 11 42 3 14 66 3 16 167 3 14 182 3 13 197 3 21 202 3 11 58 3 0 0 0 0 0 0 0 1").split()]


def find_model(name: str):
    """Return the float32 blob of a shipped model (NN section suffices) or None if not present."""
    for d in MODEL_DIRS:
        if not d:
            continue
        for fn in (name + ".nn.bin", name + ".bin"):
            p = os.path.join(d, fn)
            if os.path.exists(p):
                return np.fromfile(p, dtype=np.float32)
    return None


def synth_ids(rng, n, vocab=219):
    """Synthetic phoneme sequence in the CHS id scheme of BASELINE config 5:
    [0] + (initial in 8..28, final in 29..218, 3) * m + [0,0,0,1]."""
    m = max((n - 5) // 3, 0)
    ids = [0]
    for _ in range(m):
        ids += [int(rng.integers(8, 29)), int(rng.integers(29, min(219, vocab))), 3]
    ids += [0] * (n - 4 - 3 * m - 1 + 3) + [1]
    ids = ids[:n - 1] + [1] if len(ids) >= n else ids + [1] * (n - len(ids))
    return [min(i, vocab - 1) for i in ids]


def rel_err(a, b):
    """max|a-b| / max|b| — the waveform criterion of BASELINE.json (rel-err <= 1e-3)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def lsb_diff(a, b):
    return int(np.abs(np.asarray(a, np.int64) - np.asarray(b, np.int64)).max()) if len(a) else 0
