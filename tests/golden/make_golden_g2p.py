"""Generate tests/golden/g2p.npz from the COMPILED, UNMODIFIED reference g2p object (oracle/_ref/libstts_ref.so:
/root/reference/src/engipa/EnglishText2Id.cpp driven by oracle/ref_g2p.cpp).  Run in the build container:

    python tests/golden/make_golden_g2p.py

Contents:
  synth_*   seeded synthetic GRU section (regenerated at test time from `synth_seed`; checksum stored), pseudo-words, and the
            reference's phone ids / encoder state / first-step logits for them (its own gru() / gru_cell()).
  real_*    the shipped English model's GRU section: pseudo-words outside the reference's 125 k-word dictionary, the phone
            ids of the out-of-vocabulary branch and the IPA symbol ids of the reference's unmodified getIPAId(word).
Words whose decoder ever has a top-2 logit margin below 1e-3 are left out, so a different fp32 summation order cannot
flip an argmax in the parity tests.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import g2p_numpy as gn  # noqa: E402
from oracle import ref  # noqa: E402


def pseudo_words(seed: int, n: int, lo: int = 1, hi: int = 18, junk: float = 0.05) -> list[bytes]:
    """Seeded lower-case letter strings (a few with bytes outside a..z: digits, apostrophes, UTF-8)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        w = bytearray(int(c) for c in rng.integers(97, 123, L))
        if rng.random() < junk:
            w[int(rng.integers(0, L))] = int(rng.choice([39, 45, 48, 195, 169]))
        out.append(bytes(w))
    return out


def min_margin(m: dict, word: bytes) -> float:
    """Smallest top-2 logit gap over the decoder steps of `word` (numpy oracle)."""
    H = m["enc_w_hh"].shape[1]
    h = np.zeros(H, np.float32)
    for t in gn.letter_ids(word):
        h = gn.gru_cell(m["enc_emb"][t], h, m["enc_w_ih"], m["enc_w_hh"], m["enc_b_ih"], m["enc_b_hh"])
    tok, best = 2, np.inf
    for _ in range(gn.MAX_STEPS):
        h = gn.gru_cell(m["dec_emb"][tok], h, m["dec_w_ih"], m["dec_w_hh"], m["dec_b_ih"], m["dec_b_hh"])
        lg = np.sort(m["fc_w"] @ h + m["fc_b"])
        best = min(best, float(lg[-1] - lg[-2]))
        tok = int(np.argmax(m["fc_w"] @ h + m["fc_b"]))
        if tok == 3:
            break
    return best


def pack(words):
    return np.frombuffer(b"".join(words), np.uint8).copy(), np.cumsum([0] + [len(w) for w in words]).astype(np.int32)


def pad_preds(lists):
    a = np.full((len(lists), gn.MAX_STEPS), -1, np.int32)
    for i, p in enumerate(lists):
        a[i, :len(p)] = p
    return a


def main():
    out = {}
    # ---- synthetic section (weights scaled x4 so the decoder is not a near-uniform softmax)
    seed = 4242
    sec = gn.synthetic_section(seed, scale=4.0)
    m = gn.parse_section(sec)
    R = ref.RefG2p(sec)
    words = [w for w in pseudo_words(7, 90) if min_margin(m, w) > 1e-3][:64]
    res = [R.word(w) for w in words]
    L, O = pack(words)
    out.update(synth_seed=seed, synth_scale=4.0, synth_sha=hashlib.sha256(sec.tobytes()).hexdigest(), synth_letters=L, synth_offsets=O,
               synth_preds=pad_preds([r[0] for r in res]), synth_hidden=np.stack([r[1] for r in res]), synth_logits0=np.stack([r[2] for r in res]))
    R.close()
    # ---- shipped English model
    src = "/root/reference/models/single_speaker_english_fast.bin"
    from summertts_b200 import binfmt

    blob = np.fromfile(src, dtype=np.float32)
    sec = blob[int(binfmt.parse_model(blob)["nn_end"]):]
    m = gn.parse_section(sec)
    R = ref.RefG2p(sec)
    words, preds, ipa = [], [], []
    for w in pseudo_words(11, 400, lo=4, hi=16, junk=0.0):     # >= 4 letters: shorter unknown words are spelled out (:488-495)
        p, _, _ = R.word(w)
        ids = R.ipa_ids(w.decode())
        if ids != gn.preds_to_ipa_ids(p):      # the word is in the reference's dictionary (its IPA does not come from the GRU)
            continue
        if min_margin(m, w) <= 1e-3:
            continue
        words.append(w); preds.append(p); ipa.append(ids)
        if len(words) == 96:
            break
    L, O = pack(words)
    out.update(real_sha=hashlib.sha256(sec.tobytes()).hexdigest(), real_letters=L, real_offsets=O, real_preds=pad_preds(preds),
               real_ipa_ids=np.concatenate([np.asarray(i, np.int32) for i in ipa]),
               real_ipa_offsets=np.cumsum([0] + [len(i) for i in ipa]).astype(np.int32))
    R.close()
    np.savez_compressed(os.path.join(HERE, "g2p.npz"), **out)
    print("synthetic words", len(out["synth_offsets"]) - 1, "shipped-model words", len(words))


if __name__ == "__main__":
    main()
