"""PCM deviation statistics of the engine vs the golden reference output (shipped models) — GPU box tool."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import GOLDEN, find_model, rel_err
from summertts_b200 import engine
for name in ("single_speaker_fast", "single_speaker_mid", "multi_speakers", "single_speaker_english_fast"):
    blob = find_model(name)
    if blob is None:
        continue
    g = np.load(os.path.join(GOLDEN, "real_%s.npz" % name))
    for tp in (1, 0):
        E = engine.SynthesizerTrn(blob)
        E.set_tensor_path(tp)
        E.debug_enable(True)
        pcm = E.infer_ids(g["ids"], int(g["sid"]), float(g["ls"]))
        d = np.abs(pcm.astype(np.int64) - g["pcm"].astype(np.int64))
        print("%-28s tensor=%d  n=%d  max=%d  >1LSB=%d  >2LSB=%d  mean=%.4f  float rel=%.2e" % (
            name, tp, d.size, d.max(), int((d > 1).sum()), int((d > 2).sum()), d.mean(), rel_err(E.debug_fetch("o"), g["o"])))
        E.close()
