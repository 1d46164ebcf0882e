#!/usr/bin/env python3
"""Torch-free probe (GPU box): one synthetic batch through the engine; saves the PCM so that variants (STTS_PC_FUSED / STTS_TILE_TMA /
--tensor) can be compared sample by sample.  Run each variant under `timeout -s KILL`: a protocol bug hangs the kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from summertts_b200 import binfmt
from summertts_b200.engine import SynthesizerTrn
from parity_util import synth_ids

name = sys.argv[1] if len(sys.argv) > 1 else "single_speaker_fast"
out = sys.argv[2] if len(sys.argv) > 2 else None
B, n_ids = int(os.environ.get("PROBE_B", "24")), int(os.environ.get("PROBE_IDS", "128"))
mode = int(os.environ.get("PROBE_TENSOR", "1"))
blob = binfmt.synthetic_model(seed=11, **binfmt.ARCH[name])
E = SynthesizerTrn(blob)
E.set_tensor_path(mode)
vocab = binfmt.ARCH[name].get("vocab", 219)
ids = [synth_ids(np.random.default_rng(1234 + i), n_ids - (i % 5) * 7, vocab) for i in range(B)]
E.set_forced_durations(np.full(sum(len(x) for x in ids), 5.0, np.float32))
t0 = time.time()
for _ in range(3):
    pcm = E.infer_batch(ids)
dt = (time.time() - t0) / 3
cat = np.concatenate(pcm).astype(np.int32)
print("probe %s tensor=%d PC=%s TMA=%s: %d samples, %.1f ms/batch, fallbacks %s, checksum %d" % (
    name, mode, os.environ.get("STTS_PC_FUSED", "-"), os.environ.get("STTS_TILE_TMA", "-"), cat.size, dt * 1e3, E.tensor_fallbacks(), int(np.abs(cat).sum())))
if out:
    np.save(out, cat)
ref = os.environ.get("PROBE_REF")
if ref and os.path.exists(ref):
    r = np.load(ref)
    if r.size != cat.size:
        print("  SIZE MISMATCH vs", ref, r.size, cat.size); sys.exit(3)
    d = np.abs(r - cat)
    rel = float(np.sqrt((d.astype(np.float64) ** 2).sum() / max(1.0, (r.astype(np.float64) ** 2).sum())))
    print("  vs %s: max |diff| %d LSB, rel err %.2e" % (os.path.basename(ref), int(d.max()), rel))
