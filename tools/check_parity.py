"""Stage-by-stage parity report: CUDA engine vs the compiled reference (oracle/_ref). GPU box tool."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity_util import TEST_TXT_IDS, find_model, lsb_diff, rel_err, synth_ids  # noqa: E402

from oracle import ref  # noqa: E402
from summertts_b200 import binfmt, engine  # noqa: E402


def report(name, blob, ids, sid=0, ls=1.0, forced=None, tensor=None):
    R = ref.RefModel(blob)
    r = R.infer(ids, sid=sid, length_scale=ls, forced_w=forced)
    E = engine.SynthesizerTrn(blob)
    if tensor is not None:
        E.set_tensor_path(tensor)
    E.debug_enable(True)
    E.set_forced_durations(forced)
    t = time.time()
    pcm = E.infer_ids(ids, sid, ls)
    dt = time.time() - t
    print("== %s  T=%d F(ref)=%d S(ref)=%d S(gpu)=%d first-call %.1f ms  timing %s launches %d" % (
        name, len(ids), r.F, r.S, pcm.size, dt * 1e3, {k: round(v, 3) for k, v in E.last_timing().items()},
        E.kernel_launches()))
    ok = True
    for st in ("xx", "m", "logw", "w_ceil", "z_p", "z", "o"):
        try:
            g = E.debug_fetch(st)
        except Exception as e:  # noqa: BLE001
            print("   %-6s fetch failed: %s" % (st, e))
            ok = False
            continue
        w = getattr(r, st)
        if g.shape != w.shape:
            print("   %-6s SHAPE gpu %s ref %s" % (st, g.shape, w.shape))
            ok = False
            continue
        if st == "w_ceil":
            print("   %-6s equal=%s" % (st, np.array_equal(g, w)))
            ok &= bool(np.array_equal(g, w))
        else:
            e = rel_err(g, w)
            print("   %-6s rel_err=%.3e nan=%d" % (st, e, int(np.isnan(g).sum())))
            ok &= e < 1e-3
    if pcm.size == r.pcm.size:
        d = lsb_diff(pcm, r.pcm)
        print("   pcm    max LSB diff=%d  (>1 LSB on %d samples)" % (d, int((np.abs(pcm.astype(int) - r.pcm.astype(int)) > 1).sum())))
    else:
        ok = False
    E.close()
    R.close()
    return ok


def main():
    ref.set_threads(1)
    rng = np.random.default_rng(1234)
    allok = True
    for name, sid, ls in (("single_speaker_fast", 0, 1.0), ("multi_speakers", 10, 1.1), ("single_speaker_mid", 0, 1.0)):
        blob = find_model(name)
        if blob is None:
            print("model %s not available" % name)
            continue
        allok &= report(name, blob, TEST_TXT_IDS, sid, ls)
    for dt, dp, ms in ((1, 1, 0), (0, 0, 1), (2, 1, 0), (3, 1, 0), (0, 1, 1)):
        hp = dict(decType=dt, durPredType=dp, isMS=ms, nLayers=2, preCh=32)
        if ms:
            hp.update(spkNum=5, gin=32)
        if dt == 0:
            hp.update(upRates=(4, 2, 2), upK=(8, 4, 4), preCh=32)
        blob = binfmt.synthetic_model(seed=7 + dt, **hp)
        ids = synth_ids(rng, 23)
        forced = rng.integers(1, 5, size=len(ids)).astype(np.float32)
        allok &= report("synthetic dec%d dp%d ms%d" % (dt, dp, ms), blob, ids, sid=3, ls=1.0, forced=forced)
        allok &= report("synthetic dec%d dp%d ms%d (model durations)" % (dt, dp, ms), blob, ids, sid=3, ls=1.0)
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
