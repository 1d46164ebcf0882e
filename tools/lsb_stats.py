"""PCM / waveform deviation of the engine vs the golden reference output (shipped models), per arithmetic mode — GPU box tool.
tensor=1: split-fp16 tcgen05 (fp32-accurate, default); tensor=2: throughput mode (one fp16 MMA per K-step in the frame-level
layers); tensor=0: fp32 FFMA tiles.  Columns: int16 PCM deviation, max|a-b|/max|b| of the float waveform (BASELINE.json's
criterion), waveform SNR, and spectral convergence ||(|A|-|B|)||_F / ||(|B|)||_F of a 512/128 STFT (phase-insensitive)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import GOLDEN, find_model, rel_err
from summertts_b200 import engine


def stft_mag(x, n=512, hop=128):
    x = np.asarray(x, np.float64)
    if x.size < n:
        x = np.pad(x, (0, n - x.size))
    w = np.hanning(n)
    idx = np.arange(0, x.size - n + 1, hop)[:, None] + np.arange(n)[None, :]
    return np.abs(np.fft.rfft(x[idx] * w, axis=1))


for name in ("single_speaker_fast", "single_speaker_mid", "multi_speakers", "single_speaker_english_fast"):
    blob = find_model(name)
    if blob is None:
        continue
    g = np.load(os.path.join(GOLDEN, "real_%s.npz" % name))
    for tp in (1, 2, 0):
        E = engine.SynthesizerTrn(blob)
        E.set_tensor_path(tp)
        E.debug_enable(True)
        pcm = E.infer_ids(g["ids"], int(g["sid"]), float(g["ls"]))
        same_frames = bool(np.array_equal(E.debug_fetch("w_ceil"), g["wceil"]))
        if pcm.size != g["pcm"].size:
            print("%-28s tensor=%d  FRAME COUNT DIFFERS: %d vs %d samples" % (name, tp, pcm.size, g["pcm"].size))
            E.close()
            continue
        o = E.debug_fetch("o").ravel().astype(np.float64)
        ref = g["o"].ravel().astype(np.float64)
        d = np.abs(pcm.astype(np.int64) - g["pcm"].astype(np.int64))
        snr = 10 * np.log10((ref ** 2).sum() / max(((o - ref) ** 2).sum(), 1e-30))
        A, B = stft_mag(o), stft_mag(ref)
        sc = np.linalg.norm(A - B) / np.linalg.norm(B)
        print("%-28s tensor=%d  w_ceil %s  n=%d  max=%d  >1LSB=%d  >2LSB=%d  mean=%.4f  float rel=%.2e  SNR=%.1f dB  spec.conv=%.2e" % (
            name, tp, "exact" if same_frames else "DIFFERS", d.size, d.max(), int((d > 1).sum()), int((d > 2).sum()), d.mean(),
            rel_err(o, ref), snr, sc))
        E.close()
