"""GPU-box tool: one stts_g2p_predict launch per kernel variant on the shipped (else seeded) GRU, checked against the numpy
restatement — the target of the ncu / compute-sanitizer runs of tools/r2_final2.sh.   usage: python tools/g2p_probe.py [kernel] [n_words]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 592
os.environ["STTS_G2P_KERNEL"] = str(kernel)
from summertts_b200 import engine  # noqa: E402
from test_g2p import _real_section  # noqa: E402

from oracle import g2p_numpy as gn  # noqa: E402  (checker only)

sec = _real_section()
if sec is None:
    sec = gn.synthetic_section(4242, scale=4.0)
rng = np.random.default_rng(0)
words = [bytes(int(c) for c in rng.integers(97, 123, int(rng.integers(4, 13)))) for _ in range(n)]
g = engine.G2p(sec)
assert g.kernel == kernel, (g.kernel, g.clusters)
got = g.predict(words)
m = gn.parse_section(sec)
bad = sum(got[i] != gn.predict_word(m, w)[0] for i, w in enumerate(words[:64]))
print("g2p probe: kernel %d (clusters %d), %d words, %d of the first 64 differ from the restatement" % (g.kernel, g.clusters, n, bad))
sys.exit(1 if bad else 0)
