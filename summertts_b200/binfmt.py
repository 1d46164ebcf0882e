"""SummerTTS `.bin` model format — NN section reader and writer (host side, numpy only).

The format is the flat float32 stream consumed sequentially by the reference constructors
(SURVEY.md §8a-fmt).  Integers are stored as floats.  Record grammar and the reference ctor that
reads each record:

  file      := isMS langType durPredType decType                  SynthesizerTrn.cpp:103-106
               TextEncoder Generator<decType> Flow DurPred<durPredType>
               [spkNum gin emg(spkNum x gin, col-major)]           SynthesizerTrn.cpp:155-163
               <frontend tail: not parsed here>
  conv1d    := outCh inCh k pad dil hasBias W[outCh][k][inCh] [b]   nn_conv1d.cpp:25-52
  convT     := outCh inCh k pad dil hasBias stride W[o][k][c] [b]   nn_conv1d_transposed.cpp:25-52
  lnorm     := size gamma[size] beta[size]                          nn_layer_norm.cpp:18-34

`parse_model()` returns a nested dict of numpy views; `ModelWriter`/`synthetic_model()` emit the
same stream from a hyper-parameter dict with seeded random weights (used for the decoder variants
no shipped model reaches, for CPU-only tests and for bench runs without the shipped weights).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------------------------
class _Cur:
    def __init__(self, blob):
        self.b = blob
        self.o = 0

    def i(self):
        v = int(self.b[self.o])
        self.o += 1
        return v

    def f(self, n):
        v = self.b[self.o:self.o + n]
        if v.size != n:
            raise ValueError("model blob truncated at float offset %d" % self.o)
        self.o += n
        return v


def _conv1d(c: _Cur):
    outCh, inCh, k, pad, dil, hasBias = (c.i() for _ in range(6))
    w = c.f(outCh * k * inCh).reshape(outCh, k, inCh)
    b = c.f(outCh) if hasBias == 1 else None
    return dict(kind="conv1d", outCh=outCh, inCh=inCh, k=k, pad=pad, dil=dil, hasBias=hasBias, w=w, b=b)


def _convT(c: _Cur):
    outCh, inCh, k, pad, dil, hasBias, stride = (c.i() for _ in range(7))
    w = c.f(outCh * k * inCh).reshape(outCh, k, inCh)
    b = c.f(outCh) if hasBias == 1 else None
    return dict(kind="convT", outCh=outCh, inCh=inCh, k=k, pad=pad, dil=dil, hasBias=hasBias, stride=stride,
                w=w, b=b)


def _lnorm(c: _Cur):
    n = c.i()
    return dict(kind="lnorm", size=n, gamma=c.f(n), beta=c.f(n))


def _mha(c: _Cur):
    d = dict(channels=c.i(), outCh=c.i(), nHeads=c.i(), winSize=c.i())
    if d["winSize"] != 0:  # multi_head_attention.cpp:47-60; Map(rows, cols) col-major
        r, cc = c.i(), c.i()
        d["embRelK"] = c.f(r * cc).reshape(cc, r).T  # -> [rows][cols]
        r, cc = c.i(), c.i()
        d["embRelV"] = c.f(r * cc).reshape(cc, r).T
    for n in ("q", "k", "v", "o"):
        d[n] = _conv1d(c)
    return d


def _ffn(c: _Cur):
    return dict(k=c.i(), conv1=_conv1d(c), conv2=_conv1d(c))


def _text_encoder(c: _Cur):
    hidden, vocab, emb = c.i(), c.i(), c.i()
    E = c.f(vocab * emb).reshape(emb, vocab).T  # TextEncoder.cpp:36-38 col-major Map(vocab, emb)
    n = c.i()  # attention_encoder.cpp:30-55: all mha, all norm1, all ffn, all norm2
    mha = [_mha(c) for _ in range(n)]
    n1 = [_lnorm(c) for _ in range(n)]
    ffn = [_ffn(c) for _ in range(n)]
    n2 = [_lnorm(c) for _ in range(n)]
    proj = _conv1d(c)
    return dict(hidden=hidden, vocab=vocab, embDim=emb, emb=E, nLayers=n, mha=mha, norm1=n1, ffn=ffn, norm2=n2,
                proj=proj)


def _resblock1(c: _Cur):
    n = c.i()
    return dict(n=n, convs1=[_conv1d(c) for _ in range(n)], convs2=[_conv1d(c) for _ in range(n)])


def _generator(c: _Cur, decType: int, isMS: int):
    g = dict(decType=decType)
    if decType >= 1:  # Generator_MS.cpp:62-64
        g["subBands"], g["nfft"], g["hop"] = c.i(), c.i(), c.i()
    nUp = c.i()
    g["upRates"] = [c.i() for _ in range(nUp)]
    g["upInitCh"] = c.i()
    nUpK = c.i()
    g["upK"] = [c.i() for _ in range(nUpK)]
    nRbK = c.i()
    g["rbK"] = [c.i() for _ in range(nRbK)]
    nRbD = c.i()
    g["rbD"] = [[c.i(), c.i(), c.i()] for _ in range(nRbD)]
    g["conv_pre"] = _conv1d(c)
    g["ups"] = []
    for i in range(nUp):
        u = _convT(c)
        # ctor override: stride=u, padding=floor((k-u)/2)  (Generator_MS.cpp:88-93)
        u["stride"] = g["upRates"][i]
        u["pad"] = (g["upK"][i] - g["upRates"][i]) // 2
        g["ups"].append(u)
    g["resblocks"] = [_resblock1(c) for _ in range(nUp * nRbK)]
    if decType == 0:
        g["conv_post"] = _conv1d(c)
        if isMS == 1:
            g["cond"] = _conv1d(c)
    else:
        g["subband_conv_post"] = _conv1d(c)
        if decType == 1:
            g["multistream_conv_post"] = _conv1d(c)
    return g


def _wn(c: _Cur, isMS: int):
    n, k = c.i(), c.i()
    ins = []
    for _ in range(n):
        cv = _conv1d(c)
        cv["dil"] = 1  # WN.cpp:36-42 with dilation_rate == 1
        cv["pad"] = (k * 1 - 1) // 2
        ins.append(cv)
    rs = [_conv1d(c) for _ in range(n)]
    d = dict(nLayers=n, k=k, in_layers=ins, res_skip=rs)
    if isMS == 1:
        d["cond"] = _conv1d(c)
    return d


def _flow(c: _Cur, isMS: int):
    nFlows, nLayers = c.i(), c.i()
    layers = []
    for _ in range(nFlows):  # ResidualCouplingLayer.cpp:28-30
        layers.append(dict(pre=_conv1d(c), wn=_wn(c, isMS), post=_conv1d(c)))
    return dict(nFlows=nFlows, nLayers=nLayers, layers=layers)


def _dds(c: _Cur):
    n, k = c.i(), c.i()
    sep = []
    dil = 1
    for _ in range(n):  # DDSConv.cpp:33-41 (override pad/dil, depthwise)
        cv = _conv1d(c)
        cv["dil"] = dil
        cv["pad"] = (k * dil - dil) // 2
        cv["sep"] = 1
        sep.append(cv)
        dil *= k
    c11 = [_conv1d(c) for _ in range(n)]
    n1 = [_lnorm(c) for _ in range(n)]
    n2 = [_lnorm(c) for _ in range(n)]
    return dict(nLayers=n, k=k, sep=sep, c11=c11, norm1=n1, norm2=n2)


def _ea(c: _Cur):
    return dict(m=c.f(2), logs=c.f(2))


def _convflow(c: _Cur):
    return dict(pre=_conv1d(c), dds=_dds(c), proj=_conv1d(c))


def _dur_pred(c: _Cur, durPredType: int, isMS: int):
    if durPredType == 1:  # FixDurationPredictor.cpp:33-44
        d = dict(type=1, conv_1=_conv1d(c), norm_1=_lnorm(c), conv_2=_conv1d(c), norm_2=_lnorm(c), proj=_conv1d(c))
        if isMS == 1:
            d["cond"] = _conv1d(c)
        return d
    d = dict(type=0)  # StochasticDurationPredictor.cpp:41-70
    d["nFlows"] = c.i()
    d["ea"] = _ea(c)
    d["flows"] = [_convflow(c) for _ in range(d["nFlows"])]
    d["post_pre"] = _conv1d(c)
    d["post_proj"] = _conv1d(c)
    d["post_convs"] = _dds(c)
    d["post_ea"] = _ea(c)
    d["post_flows"] = [_convflow(c) for _ in range(4)]
    d["pre"] = _conv1d(c)
    d["proj"] = _conv1d(c)
    d["convs"] = _dds(c)
    if isMS == 1:
        d["cond"] = _conv1d(c)
    return d


def parse_model(blob: np.ndarray) -> dict:
    """Parse the NN section of a SummerTTS .bin (float32 array). Returns the layer tree + nn_end."""
    blob = np.asarray(blob, dtype=np.float32)
    c = _Cur(blob)
    M = dict(isMS=c.i(), langType=c.i(), durPredType=c.i(), decType=c.i())
    M["enc"] = _text_encoder(c)
    M["dec"] = _generator(c, M["decType"], M["isMS"])
    M["flow"] = _flow(c, M["isMS"])
    M["dp"] = _dur_pred(c, M["durPredType"], M["isMS"])
    M["spkNum"], M["gin"] = 0, 0
    if M["isMS"] == 1:
        M["spkNum"], M["gin"] = c.i(), c.i()
        M["emg"] = c.f(M["spkNum"] * M["gin"]).reshape(M["gin"], M["spkNum"]).T  # [spk][gin]
    M["nn_end"] = c.o
    return M


# ----------------------------------------------------------------------------------------------
# writer (synthetic models)
# ----------------------------------------------------------------------------------------------
class ModelWriter:
    def __init__(self, seed=0, wscale=1.0):
        self.rng = np.random.default_rng(seed)
        self.parts = []
        self.wscale = wscale

    def ints(self, *v):
        self.parts.append(np.asarray(v, dtype=np.float32))

    def floats(self, a):
        self.parts.append(np.asarray(a, dtype=np.float32).ravel())

    def conv1d(self, outCh, inCh, k, pad=0, dil=1, bias=True, gain=1.0, fan_in=None):
        self.ints(outCh, inCh, k, pad, dil, 1 if bias else 0)
        fan = fan_in if fan_in is not None else inCh * k
        self.floats(self.rng.standard_normal((outCh, k, inCh)) * (gain * self.wscale / np.sqrt(fan)))
        if bias:
            self.floats(self.rng.standard_normal(outCh) * 0.05)

    def convT(self, outCh, inCh, k, stride, pad, bias=True, gain=1.0):
        self.ints(outCh, inCh, k, pad, 1, 1 if bias else 0, stride)
        self.floats(self.rng.standard_normal((outCh, k, inCh)) * (gain * self.wscale / np.sqrt(inCh * k / stride)))
        if bias:
            self.floats(self.rng.standard_normal(outCh) * 0.05)

    def lnorm(self, n):
        self.ints(n)
        self.floats(1.0 + 0.1 * self.rng.standard_normal(n))
        self.floats(0.1 * self.rng.standard_normal(n))

    def blob(self):
        return np.concatenate(self.parts).astype(np.float32)


DEFAULT_HP = dict(
    isMS=0, langType=0, durPredType=1, decType=1,
    hidden=192, vocab=219, nLayers=6, ffn=768, ffnK=3, nHeads=2, winSize=4,
    inter=192, flowN=4, wnLayers=4, wnK=5, wnHidden=192,
    dpFilter=256, dpK=3, sdpFlows=4, sdpK=3, sdpLayers=3,
    upRates=(4, 4), upK=(16, 16), preCh=128, rbK=(3, 7, 11), rbD=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
    spkNum=0, gin=0,
)

# architecture of the shipped files (SURVEY.md §8a-fmt "Shipped hyper-parameters")
ARCH = {
    "single_speaker_fast": dict(),
    "single_speaker_english_fast": dict(langType=1, vocab=178),
    "single_speaker_mid": dict(preCh=256),
    "single_speaker_english": dict(langType=1, vocab=178, preCh=256),
    "multi_speakers": dict(isMS=1, durPredType=0, decType=0, upRates=(8, 8, 2, 2), upK=(16, 16, 4, 4), preCh=64,
                           spkNum=218, gin=256),
}


def synthetic_model(seed=0, **over) -> np.ndarray:
    """Seeded random-weight model with the reference's serialisation order. Returns float32 blob."""
    hp = dict(DEFAULT_HP)
    hp.update(over)
    W = ModelWriter(seed)
    H, isMS = hp["hidden"], hp["isMS"]
    kc = H // hp["nHeads"]
    W.ints(isMS, hp["langType"], hp["durPredType"], hp["decType"])
    # --- TextEncoder
    W.ints(H, hp["vocab"], H)
    W.floats(W.rng.standard_normal((H, hp["vocab"])) * (H ** -0.5))  # stored [emb][vocab]
    W.ints(hp["nLayers"])
    for _ in range(hp["nLayers"]):
        W.ints(H, H, hp["nHeads"], hp["winSize"])
        if hp["winSize"]:
            r = 2 * hp["winSize"] + 1
            for _e in range(2):
                W.ints(r, kc)
                W.floats(W.rng.standard_normal((kc, r)) * (kc ** -0.5))
        for _c in range(4):
            W.conv1d(H, H, 1)
    for _ in range(hp["nLayers"]):
        W.lnorm(H)
    for _ in range(hp["nLayers"]):
        W.ints(hp["ffnK"])
        W.conv1d(hp["ffn"], H, hp["ffnK"])
        W.conv1d(H, hp["ffn"], hp["ffnK"])
    for _ in range(hp["nLayers"]):
        W.lnorm(H)
    W.conv1d(2 * hp["inter"], H, 1)
    # --- Generator
    dt = hp["decType"]
    if dt >= 1:
        W.ints(4, 16, 4)
    ups, upK = hp["upRates"], hp["upK"]
    W.ints(len(ups), *ups)
    W.ints(512)
    W.ints(len(upK), *upK)
    W.ints(len(hp["rbK"]), *hp["rbK"])
    W.ints(len(hp["rbD"]))
    for d3 in hp["rbD"]:
        W.ints(*d3)
    ch = hp["preCh"]
    W.conv1d(ch, hp["inter"], 7, pad=3)
    for u, k in zip(ups, upK):
        W.convT(ch // 2, ch, k, u, (k - u) // 2)
        ch //= 2
    ch = hp["preCh"]
    for _u in ups:
        ch //= 2
        for k, d3 in zip(hp["rbK"], hp["rbD"]):
            W.ints(3)
            for d in d3:
                W.conv1d(ch, ch, k, pad=(k * d - d) // 2, dil=d, gain=0.7)
            for _d in d3:
                W.conv1d(ch, ch, k, pad=(k - 1) // 2, dil=1, gain=0.5)
    if dt == 0:
        W.conv1d(1, ch, 7, pad=3, bias=False)
        if isMS:
            W.conv1d(hp["preCh"], hp["gin"], 1)
    else:
        nb = 4 if dt in (1, 3) else 1
        W.conv1d(nb * 18, ch, 7, pad=3, gain=0.5)
        if dt == 1:
            W.conv1d(1, 4, 63, pad=31, bias=False, gain=0.7)
    # --- Flow
    W.ints(hp["flowN"], hp["wnLayers"])
    half, wh = hp["inter"] // 2, hp["wnHidden"]
    for _ in range(hp["flowN"]):
        W.conv1d(wh, half, 1)
        W.ints(hp["wnLayers"], hp["wnK"])
        for _l in range(hp["wnLayers"]):
            W.conv1d(2 * wh, wh, hp["wnK"], pad=(hp["wnK"] - 1) // 2)
        for l in range(hp["wnLayers"]):
            W.conv1d(2 * wh if l < hp["wnLayers"] - 1 else wh, wh, 1, gain=0.5)
        if isMS:
            W.conv1d(2 * wh * hp["wnLayers"], hp["gin"], 1, gain=0.5)
        W.conv1d(half, wh, 1, gain=0.5)
    # --- duration predictor
    if hp["durPredType"] == 1:
        F = hp["dpFilter"]
        W.conv1d(F, H, hp["dpK"], pad=hp["dpK"] // 2)
        W.lnorm(F)
        W.conv1d(F, F, hp["dpK"], pad=hp["dpK"] // 2)
        W.lnorm(F)
        W.conv1d(1, F, 1)
        if isMS:
            W.conv1d(H, hp["gin"], 1)
    else:
        F = H

        def dds():
            W.ints(hp["sdpLayers"], hp["sdpK"])
            for _l in range(hp["sdpLayers"]):
                W.conv1d(F, 1, hp["sdpK"], pad=0, fan_in=hp["sdpK"])  # depthwise: inCh field = 1
            for _l in range(hp["sdpLayers"]):
                W.conv1d(F, F, 1)
            for _l in range(2 * hp["sdpLayers"]):
                W.lnorm(F)

        def convflow():
            W.conv1d(F, 1, 1)
            dds()
            W.conv1d(29, F, 1, gain=2.0)

        def ea():
            W.floats(0.3 * W.rng.standard_normal(2))
            W.floats(0.2 * W.rng.standard_normal(2))

        W.ints(hp["sdpFlows"])
        ea()
        for _ in range(hp["sdpFlows"]):
            convflow()
        W.conv1d(F, 1, 1)
        W.conv1d(F, F, 1)
        dds()
        ea()
        for _ in range(4):
            convflow()
        W.conv1d(F, H, 1)
        W.conv1d(F, F, 1)
        dds()
        if isMS:
            W.conv1d(F, hp["gin"], 1)
    if isMS:
        W.ints(hp["spkNum"], hp["gin"])
        W.floats(W.rng.standard_normal((hp["gin"], hp["spkNum"])) * 0.5)  # col-major (spk, gin)
    return W.blob()
