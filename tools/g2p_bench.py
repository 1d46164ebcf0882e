"""GPU-box tool: words per second of the batched GRU g2p (stts_g2p_predict, host buffers in and out, one launch per batch)
next to the reference's per-word Eigen GRU (compiled unmodified gru()/gru_cell() through oracle/ref_g2p.cpp, one thread).
usage: python tools/g2p_bench.py [n_words]  ->  one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from summertts_b200 import engine  # noqa: E402
from test_g2p import _real_section  # noqa: E402

from oracle import g2p_numpy as gn  # noqa: E402  (checker + CPU baseline only)
from oracle import ref  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    sec = _real_section()
    weights = "shipped single_speaker_english_fast"
    if sec is None:
        sec, weights = gn.synthetic_section(4242, scale=4.0), "synthetic seed 4242"
    rng = np.random.default_rng(0)
    words = [bytes(int(c) for c in rng.integers(97, 123, int(rng.integers(4, 13)))) for _ in range(n)]
    out = {"weights": weights, "n_words": n, "letters_per_word": "4..12"}
    got = None
    for kernel, name in ((0, "stream"), (1, "cluster"), (2, "auto")):     # STTS_G2P_KERNEL: W_hh streamed from L2 per step / resident in 8-CTA clusters
        os.environ["STTS_G2P_KERNEL"] = str(kernel)
        try:
            g = engine.G2p(sec)
        except engine.SttsError as e:
            out[name] = {"error": str(e)}
            continue
        if g.kernel != kernel:
            out[name] = {"error": "not selected (clusters admitted: %d)" % g.clusters}
            continue
        res = {"clusters": g.clusters} if kernel else {}
        for batch in (1, 64, 144, 592, 2048, n):
            g.predict(words[:batch])
            best = 1e9
            for _ in range(5):
                t = time.perf_counter(); r = g.predict(words[:batch]); best = min(best, time.perf_counter() - t)
            res["batch_%d" % batch] = {"ms": round(best * 1e3, 3), "words_per_s": round(batch / best)}
        if got is not None:
            res["ids_equal_to_stream"] = bool(r == got)
        got = got if got is not None else r
        out[name] = res
        g.close()
    if ref.available():
        ref.set_threads(1)
        R = ref.RefG2p(sec)
        k = 256
        t = time.perf_counter()
        want = [R.word(w)[0] for w in words[:k]]
        dt = time.perf_counter() - t
        out["reference_cpu_1_thread"] = {"ms_per_word": round(dt / k * 1e3, 4), "words_per_s": round(k / dt)}
        out["ids_equal_on_sample"] = bool(want == got[:k])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
