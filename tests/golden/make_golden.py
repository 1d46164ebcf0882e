"""Generate the committed golden fixtures under tests/golden/ from the COMPILED, UNMODIFIED reference
(oracle/_ref/libstts_ref.so, built by oracle/Makefile from /root/reference).  Run in the build
container (needs /root/reference for the shipped models); the fixtures then travel with the repo.

    python tests/golden/make_golden.py

Fixtures:
  synth_<tag>.npz   seeded synthetic model (regenerated at test time from `hp`/`seed`; blob checksum
                    stored), ids, sid, forced durations, and the reference's w_ceil / F / o / pcm /
                    stage checksums for (a) model-predicted and (b) forced durations.
  real_<model>.npz  shipped model + ids -> w_ceil, F, pcm, float waveform (float16-free: o stored as
                    float32) — only checked when the model file is available at test time.
  ops.npz           op-level known answers (conv1d dense/dilated/depthwise, ConvTranspose1d,
                    LayerNorm, iSTFT, PQMF, tanh, gelu) with their inputs.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity_util import TEST_TXT_IDS, find_model, synth_ids  # noqa: E402

from oracle import ref  # noqa: E402
from summertts_b200 import binfmt  # noqa: E402

SYNTH = {
    "ms_fix": dict(decType=1, durPredType=1, isMS=0, nLayers=2, preCh=32),
    "hifigan_sdp_spk": dict(decType=0, durPredType=0, isMS=1, nLayers=2, preCh=32, spkNum=5, gin=32,
                            upRates=(4, 2, 2), upK=(8, 4, 4)),
    "istft_fix": dict(decType=2, durPredType=1, isMS=0, nLayers=2, preCh=32),
    "mbb_fix": dict(decType=3, durPredType=1, isMS=0, nLayers=2, preCh=32),
    "hifigan_fix_spk": dict(decType=0, durPredType=1, isMS=1, nLayers=2, preCh=32, spkNum=5, gin=32,
                            upRates=(4, 2, 2), upK=(8, 4, 4)),
}


def blob_sum(blob):
    return hashlib.sha256(np.ascontiguousarray(blob, dtype=np.float32).tobytes()).hexdigest()


def main():
    ref.set_threads(1)
    rng = np.random.default_rng(20260923)
    for i, (tag, hp) in enumerate(SYNTH.items()):
        seed = 100 + i
        blob = binfmt.synthetic_model(seed=seed, **hp)
        R = ref.RefModel(blob)
        ids = np.array(synth_ids(rng, 23 + 2 * i), dtype=np.int32)
        forced = rng.integers(0, 5, size=ids.size).astype(np.float32)
        forced[0] = 2
        a = R.infer(ids, sid=3, length_scale=1.0)
        b = R.infer(ids, sid=3, length_scale=1.0, forced_w=forced)
        np.savez_compressed(
            os.path.join(HERE, "synth_%s.npz" % tag), hp=json.dumps(hp), seed=seed, sha=blob_sum(blob), ids=ids,
            sid=3, forced=forced,
            a_wceil=a.w_ceil, a_F=a.F, a_o=a.o, a_pcm=a.pcm, a_xx=a.xx, a_m=a.m, a_logw=a.logw, a_z=a.z,
            b_wceil=b.w_ceil, b_F=b.F, b_o=b.o, b_pcm=b.pcm, b_z=b.z)
        print(tag, "F", a.F, b.F, "S", a.S, b.S)
        R.close()
    for name, sid, ls in (("single_speaker_fast", 0, 1.0), ("single_speaker_mid", 0, 1.0), ("multi_speakers", 10, 1.1),
                          ("single_speaker_english_fast", 0, 0.83)):
        blob = find_model(name)
        if blob is None:
            print("skip", name)
            continue
        M = binfmt.parse_model(blob)
        # CHS models: the ids of /root/reference/test.txt; ENG: the 700 ids of /root/reference/test_eng.txt (BASELINE config 4),
        # both through the reference's own text frontend (summertts_b200/host/_build/tts_b200 --dump-ids, committed as
        # tests/golden/test_eng_ids.txt)
        if M["langType"] == 0:
            ids = np.array(TEST_TXT_IDS, dtype=np.int32)
        else:
            ids = np.array(open(os.path.join(HERE, "test_eng_ids.txt")).read().split(), dtype=np.int32)
        R = ref.RefModel(blob)
        r = R.infer(ids, sid=sid, length_scale=ls)
        np.savez_compressed(os.path.join(HERE, "real_%s.npz" % name), ids=ids, sid=sid, ls=np.float32(ls),
                            wceil=r.w_ceil, F=r.F, pcm=r.pcm, o=r.o.astype(np.float32), nn_end=M["nn_end"],
                            sha=blob_sum(blob[:M["nn_end"]]))
        print(name, "F", r.F, "S", r.S)
        R.close()
    # ---- op-level KATs ----------------------------------------------------------------------------
    ops = {}
    x = rng.standard_normal((37, 12)).astype(np.float32)

    def rec_conv(o, c, k, p, d, bias=True):
        w = rng.standard_normal((o, k, c)).astype(np.float32)
        b = rng.standard_normal(o).astype(np.float32)
        rec = np.concatenate([np.array([o, c, k, p, d, 1 if bias else 0], np.float32), w.ravel(), b if bias else []])
        return rec.astype(np.float32)

    for tag, (o, k, p, d) in {"dense": (7, 3, 1, 1), "dil": (5, 5, 6, 3), "k1": (9, 1, 0, 1)}.items():
        rec = rec_conv(o, 12, k, p, d)
        ops["conv_%s_rec" % tag] = rec
        ops["conv_%s_y" % tag] = ref.conv1d(rec, x)
    rec = rec_conv(12, 1, 3, 0, 1)
    ops["dw_rec"] = rec
    ops["dw_y"] = ref.conv1d(rec, x, mode=1, pad=3, dil=3, sep=1)
    w = rng.standard_normal((6, 8, 12)).astype(np.float32)
    b = rng.standard_normal(6).astype(np.float32)
    rec = np.concatenate([np.array([6, 12, 8, 0, 1, 1, 1], np.float32), w.ravel(), b]).astype(np.float32)
    ops["convT_rec"] = rec
    ops["convT_y"] = ref.conv1d_transposed(rec, x, 4, 2)
    rec = np.concatenate([[12], rng.standard_normal(12), rng.standard_normal(12)]).astype(np.float32)
    ops["ln_rec"] = rec
    ops["ln_y"] = ref.layer_norm(rec, x)
    mag = np.exp(0.3 * rng.standard_normal((21, 9))).astype(np.float32)
    ph = (np.pi * np.sin(rng.standard_normal((21, 9)))).astype(np.float32)
    ops["istft_mag"], ops["istft_ph"] = mag, ph
    ops["istft_y"] = ref.istft(mag, ph)
    xb = rng.standard_normal((40, 4)).astype(np.float32)
    ops["pqmf_x"] = xb
    ops["pqmf_y"] = ref.pqmf(xb)
    v = np.concatenate([rng.standard_normal(64) * 4, [-100, 100, 0, 90, -90]]).astype(np.float32)
    ops["elt_x"] = v
    ops["tanh_y"] = ref.eltwise(0, v)
    ops["gelu_y"] = ref.eltwise(1, v)
    ops["x"] = x
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)
    print("ops written")


if __name__ == "__main__":
    main()
