"""Error of the two conv paths against an exact (float64) convolution — GPU box tool."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from summertts_b200 import engine

rng = np.random.default_rng(0)
for (o, c, k, d) in [(64, 64, 11, 5), (384, 192, 5, 1), (192, 768, 3, 1), (32, 32, 3, 1), (192, 192, 1, 1)]:
    w = (rng.standard_normal((o, k, c)) * 0.3).astype(np.float32)
    b = rng.standard_normal(o).astype(np.float32)
    p = d * (k - 1) // 2
    rec = np.concatenate([np.array([o, c, k, p, d, 1], np.float32), w.ravel(), b]).astype(np.float32)
    T = 512
    x = (np.abs(rng.standard_normal((T, c))) * 3).astype(np.float32)  # positive inputs: coherent sums expose truncation bias
    xp = np.zeros((T + 2 * p, c)); xp[p:p + T] = x
    cols = np.concatenate([xp[kk * d:kk * d + T] for kk in range(k)], axis=1)
    exact = cols @ w.astype(np.float64).reshape(o, -1).T + b
    out = []
    for tc in (0, 1):
        y = engine.test_conv1d(rec, x, use_tc=tc)
        e = np.abs(y - exact)
        out.append("%s max %.2e mean %.2e (rel to max|y| %.1f)" % ("tc  " if tc else "ffma", e.max() / np.abs(exact).max(), e.mean() / np.abs(exact).max(), np.abs(exact).max()))
    print((o, c, k, d), " | ".join(out))
