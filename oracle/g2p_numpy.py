"""oracle/g2p_numpy.py — TEST INFRASTRUCTURE ONLY (the product never imports this).

numpy restatement of the GRU grapheme-to-phoneme branch of the reference's English frontend,
/root/reference/src/engipa/EnglishText2Id.cpp:
  parse_section   constructor record walk                         :73-126   (Eigen column-major Maps)
  gru_cell        :270-294     gru :296-313
  predict_word    out-of-vocabulary branch of getIPAId            :496-545  (encoder over letters + </s>, greedy decoder)
  preds_to_ipa_ids  phone ids -> IPA string -> symbol ids         :547-606  (constant tables :69, :160-259)
  synthetic_section seeded GRU section writer for fixtures (same record layout)

Pinned (tests/test_g2p.py, CPU) against the compiled unmodified reference: predict_word vs oracle/ref_g2p.cpp's
sref_g2p_word (the reference's own gru / gru_cell), and preds_to_ipa_ids(predict_word(w)) vs the reference's own
getIPAId(w) for words outside its dictionary; against tests/golden/g2p.npz where the reference did not travel.
"""
from __future__ import annotations

import numpy as np

MAX_STEPS = 20  # :527

# constant tables of the frontend (EnglishText2Id.cpp:69, :160-233, :235-259)
IPA_SYMBOLS = ('_;:,.!?¡¿—…"«»“” ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz'
               'ɑɐɒæɓʙβɔɕçɗɖðʤəɘɚɛɜɝɞɟʄɡɠɢʛɦɧħɥʜɨɪʝɭɬɫɮʟɱɯɰŋɳɲɴøɵɸθœɶʘɹɺɾɻʀʁɽʂʃʈʧʉʊʋⱱʌɣɤʍχʎʏʑʐʒʔʡʕʢǀǁǂǃˈˌːˑʼʴʰʱʲʷˠˤ˞↓↑→↗↘\'̩\'ᵻ')
ID2PHONE = ['<pad>', '<unk>', '<s>', '</s>', 'AA0', 'AA1', 'AA2', 'AE0', 'AE1', 'AE2', 'AH0', 'AH1', 'AH2', 'AO0', 'AO1', 'AO2',
            'AW0', 'AW1', 'AW2', 'AY0', 'AY1', 'AY2', 'B', 'CH', 'D', 'DH', 'EH0', 'EH1', 'EH2', 'ER0', 'ER1', 'ER2', 'EY0', 'EY1',
            'EY2', 'F', 'G', 'HH', 'IH0', 'IH1', 'IH2', 'IY0', 'IY1', 'IY2', 'JH', 'K', 'L', 'M', 'N', 'NG', 'OW0', 'OW1', 'OW2',
            'OY0', 'OY1', 'OY2', 'P', 'R', 'S', 'SH', 'T', 'TH', 'UH0', 'UH1', 'UH2', 'UW', 'UW0', 'UW1', 'UW2', 'V', 'W', 'Y', 'Z', 'ZH']
PHONE2IPA = {'a': 'ə', 'ey': 'eɪ', 'aa': 'ɑ', 'ae': 'æ', 'ah': 'ə', 'ao': 'ɔ', 'aw': 'aʊ', 'ay': 'aɪ', 'ch': 'ʧ', 'dh': 'ð', 'eh': 'ɛ',
             'er': 'ər', 'hh': 'h', 'ih': 'ɪ', 'jh': 'ʤ', 'ng': 'ŋ', 'ow': 'oʊ', 'oy': 'ɔɪ', 'sh': 'ʃ', 'th': 'θ', 'uh': 'ʊ', 'uw': 'u',
             'zh': 'ʒ', 'iy': 'i', 'y': 'j'}

NAMES = ("enc_emb", "enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "dec_emb", "dec_w_ih", "dec_w_hh", "dec_b_ih", "dec_b_hh", "fc_w", "fc_b")
_IS_VEC = {"enc_b_ih", "enc_b_hh", "dec_b_ih", "dec_b_hh", "fc_b"}


def parse_section(sec: np.ndarray) -> dict:
    """:73-126.  Matrices come back as (rows, cols) arrays (the file holds them column-major)."""
    sec = np.asarray(sec, dtype=np.float32)
    cur, out = 0, {}
    for name in NAMES:
        if name in _IS_VEC:
            n = int(sec[cur]); cur += 1
            out[name] = sec[cur:cur + n].copy(); cur += n
        else:
            r, c = int(sec[cur]), int(sec[cur + 1]); cur += 2
            out[name] = sec[cur:cur + r * c].reshape(c, r).T.copy(); cur += r * c
    out["consumed"] = cur
    return out


def synthetic_section(seed: int, hidden: int = 256, emb: int = 256, n_letters: int = 29, n_phones: int = 74, scale: float = 1.0) -> np.ndarray:
    """Seeded GRU section in the reference's record layout (uniform +-scale/sqrt(hidden), PyTorch's GRU init)."""
    rng = np.random.default_rng(seed)
    k = scale / np.sqrt(hidden)
    shapes = {"enc_emb": (n_letters, emb), "enc_w_ih": (3 * hidden, emb), "enc_w_hh": (3 * hidden, hidden), "enc_b_ih": (3 * hidden,),
              "enc_b_hh": (3 * hidden,), "dec_emb": (n_phones, emb), "dec_w_ih": (3 * hidden, emb), "dec_w_hh": (3 * hidden, hidden),
              "dec_b_ih": (3 * hidden,), "dec_b_hh": (3 * hidden,), "fc_w": (n_phones, hidden), "fc_b": (n_phones,)}
    parts = []
    for name in NAMES:
        shp = shapes[name]
        if name.endswith("emb"):
            a = rng.standard_normal(shp).astype(np.float32)
        else:
            a = rng.uniform(-k, k, shp).astype(np.float32)
        parts.append(np.asarray(shp, dtype=np.float32))
        parts.append(a.T.reshape(-1) if a.ndim == 2 else a)     # column-major
    return np.concatenate(parts).astype(np.float32)


def _sigmoid(x):   # nn_sigmoid.cpp:3-7
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def _tanh(x):      # nn_tanh.cpp:6-21
    x = x.astype(np.float32)
    with np.errstate(over="ignore"):
        a, b = np.exp(x), np.exp(-x)
    a = np.where(np.isinf(a), np.float32(1e10), a)
    b = np.where(np.isinf(b), np.float32(1e10), b)
    d = np.maximum(a + b, np.float32(1e-8))
    return ((a - b) / d).astype(np.float32)


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """:270-294.  x[E], h[H] -> h'[H]."""
    H = h.shape[0]
    rzn_ih = w_ih @ x + b_ih
    rzn_hh = w_hh @ h + b_hh
    rz = _sigmoid(rzn_ih[:2 * H] + rzn_hh[:2 * H])
    r, z = rz[:H], rz[H:]
    n = _tanh(rzn_ih[2 * H:] + r * rzn_hh[2 * H:])
    return ((z * np.float32(-1) + np.float32(1)) * n + z * h).astype(np.float32)


def letter_ids(word: bytes) -> list[int]:
    """:496-513: per BYTE, 'a'..'z' -> 3..28, anything else -> <unk> = 1; then </s> = 2."""
    return [3 + (c - 97) if 97 <= c <= 122 else 1 for c in word] + [2]


def predict_word(m: dict, word: bytes | str):
    """:496-545.  Returns (preds, encoder hidden state, logits of the first decoder step)."""
    if isinstance(word, str):
        word = word.encode("utf-8")
    H = m["enc_w_hh"].shape[1]
    h = np.zeros(H, np.float32)
    for t in letter_ids(word):                                          # gru(), :296-313
        h = gru_cell(m["enc_emb"][t], h, m["enc_w_ih"], m["enc_w_hh"], m["enc_b_ih"], m["enc_b_hh"])
    hidden = h.copy()
    tok, preds, logits0 = 2, [], None
    for i in range(MAX_STEPS):
        h = gru_cell(m["dec_emb"][tok], h, m["dec_w_ih"], m["dec_w_hh"], m["dec_b_ih"], m["dec_b_hh"])
        logits = (m["fc_w"] @ h + m["fc_b"]).astype(np.float32)
        if i == 0:
            logits0 = logits.copy()
        tok = int(np.argmax(logits))                                    # first maximum, like Eigen's maxCoeff
        if tok == 3:
            break
        preds.append(tok)
    return preds, hidden, logits0


def preds_to_ipa(preds) -> str:
    """:547-566: phone name without its stress digit, lower-cased, through phone2ipa_ (else the name itself)."""
    out = []
    for p in preds:
        name = "".join(ch for ch in ID2PHONE[p] if not ch.isdigit()).lower()
        out.append(PHONE2IPA.get(name, name))
    return "".join(out)


def ipa_to_ids(ipa: str) -> list[int]:
    """:575-603 for one entry of vecIPAs: (0, k) per known symbol, 16 per unknown one, then the word separator (0, 16)."""
    ids = []
    for ch in ipa:
        k = IPA_SYMBOLS.find(ch)
        ids += [0, k] if k >= 0 else [16]
    return ids + [0, 16]


def preds_to_ipa_ids(preds) -> list[int]:
    return ipa_to_ids(preds_to_ipa(preds))
