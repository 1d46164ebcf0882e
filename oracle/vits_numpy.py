"""oracle/vits_numpy.py — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (numpy, float32) of the reference's VITS acoustic+vocoder forward pass, i.e. the NN
half of `SynthesizerTrn::infer` (/root/reference/src/models/SynthesizerTrn.cpp:357-396) and every
src/nn_op, src/modules, src/models function it reaches.  Each function cites the reference
file:line it restates.  All activations are time-major [T][C] (the reference's MatrixXf is
[T rows x C cols]).

PINNING: this restatement is checked in tests/test_oracle.py against
  (a) the compiled, unmodified reference objects (oracle/_ref/libstts_ref.so, built by
      oracle/Makefile from /root/reference in place) on op-level and whole-model inputs, and
  (b) the committed golden fixtures under tests/golden/ that were generated from (a) by
      tests/golden/make_golden.py.
The reference ships no tests/golden vectors of its own; (a) reproduces the md5 of
the reference CLI's WAV output for three shipped models, which is what pins it.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------------
# nn_op
# ----------------------------------------------------------------------------------------------
def conv1d(x, cv, pad=None, dil=None):
    """nn_conv1d::forward, src/nn_op/nn_conv1d.cpp:118-199 (dense and depthwise `sep_` branch).
    y[t,o] = b[o] + sum_k sum_c xpad[t + k*d, c] * W[o][k][c];  W is the file order [o][k][c]."""
    x = np.asarray(x, f32)
    p = cv["pad"] if pad is None else pad
    d = cv["dil"] if dil is None else dil
    k, Co = cv["k"], cv["outCh"]
    T = x.shape[0]
    xp = np.zeros((T + 2 * p, x.shape[1]), f32)
    xp[p:p + T] = x
    To = T + 2 * p - d * (k - 1)
    if cv.get("sep", 0):
        # depthwise: record has inCh == 1, w[o][k][0]   (nn_conv1d.cpp:167-179)
        y = np.zeros((To, Co), f32)
        for kk in range(k):
            y += xp[kk * d:kk * d + To, :] * cv["w"][:, kk, 0][None, :]
    else:
        cols = np.concatenate([xp[kk * d:kk * d + To, :] for kk in range(k)], axis=1)  # [To][k*Cin]
        y = cols @ cv["w"].reshape(Co, -1).T
    if cv["hasBias"] == 1:
        y = y + cv["b"][None, :]
    return y.astype(f32)


def conv1d_transposed(x, cv):
    """nn_conv1d_transposed::forward, src/nn_op/nn_conv1d_transposed.cpp:106-150.
    y[i*s + kk - p, o] += x[i,c] * W[o][kk][c]; bias; crop to outLen=(T-1)s-2p+(k-1)+1."""
    x = np.asarray(x, f32)
    s, p, k, Co = cv["stride"], cv["pad"], cv["k"], cv["outCh"]
    T = x.shape[0]
    outLen = (T - 1) * s - 2 * p + (k - 1) + 1
    full = np.zeros((outLen + 2 * p + k, Co), f32)
    contrib = x @ cv["w"].reshape(Co * k, -1).T  # [T][o*k + kk]
    contrib = contrib.reshape(T, Co, k)
    for i in range(T):
        full[i * s:i * s + k, :] += contrib[i].T
    if cv["hasBias"] == 1:
        full = full + cv["b"][None, :]
    return full[p:p + outLen].astype(f32)


def layer_norm(x, ln):
    """nn_layer_norm::forward, src/nn_op/nn_layer_norm.cpp:65-86: var = E[x^2] - mean^2, eps 1e-5."""
    x = np.asarray(x, f32)
    mean = x.mean(axis=1, keepdims=True, dtype=f32)
    sq = (x * x).sum(axis=1, keepdims=True, dtype=f32)
    var = sq * f32(1.0 / x.shape[1]) - mean * mean
    return (((x - mean) / np.sqrt(var + f32(1e-5))) * ln["gamma"][None, :] + ln["beta"][None, :]).astype(f32)


def tanh_ref(x):
    """nn_tanh, src/nn_op/nn_tanh.cpp:6-21: (e^x - e^-x)/(e^x + e^-x) with inf->1e10, denom floor 1e-8."""
    x = np.asarray(x, f32)
    with np.errstate(over="ignore"):
        a = np.exp(x)
        b = np.exp(-x)
    a = np.where(np.isinf(a), f32(1e10), a)
    b = np.where(np.isinf(b), f32(1e10), b)
    m1 = a + b
    m1 = np.where(m1 < f32(1e-8), f32(1e-8), m1)
    return ((a - b) / m1).astype(f32)


def sigmoid(x):
    """nn_sigmoid, src/nn_op/nn_sigmoid.cpp:3-7."""
    with np.errstate(over="ignore"):
        return (f32(1.0) / (f32(1.0) + np.exp(-np.asarray(x, f32)))).astype(f32)


def gelu(x):
    """nn_gelu, src/nn_op/nn_gelu.cpp:7-14 (tanh approximation through nn_tanh)."""
    x = np.asarray(x, f32)
    t = tanh_ref((x + x * x * x * f32(0.044715)) * f32(0.7978845608028654))
    return ((t + f32(1.0)) * x * f32(0.5)).astype(f32)


def leaky_relu(x, slope=0.01):
    """nn_leaky_relu, src/nn_op/nn_leaky_relu.cpp:6-27."""
    x = np.asarray(x, f32)
    return np.where(x < 0, x * f32(slope), x).astype(f32)


def relu(x):
    """nn_relu, src/nn_op/nn_relu.cpp:3-17."""
    return np.maximum(np.asarray(x, f32), f32(0))


def softmax_rows(x):
    """nn_softmax(dim=0), src/nn_op/nn_softmax.cpp:5-28: exp / row-sum, NO max subtraction."""
    e = np.exp(np.asarray(x, f32))
    return (e / e.sum(axis=1, keepdims=True, dtype=f32)).astype(f32)


def softplus(x):
    """nn_softplus, src/nn_op/nn_softplus.cpp:3-8: log(1 + e^x), no threshold."""
    return np.log(np.exp(np.asarray(x, f32)) + f32(1.0)).astype(f32)


# ----------------------------------------------------------------------------------------------
# text encoder
# ----------------------------------------------------------------------------------------------
def mha(x, L):
    """multi_head_attention::forward/attention, src/modules/multi_head_attention.cpp:106-121,201-295.
    score[i,j] = (q_i/sqrt(kc)).k_j + [|j-i|<=w] (q_i/sqrt(kc)).Ek[j-i+w]; p = softmax_j (no max);
    out_i = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w]; heads concatenated."""
    T = x.shape[0]
    H, w = L["nHeads"], L["winSize"]
    q, k, v = conv1d(x, L["q"]), conv1d(x, L["k"]), conv1d(x, L["v"])
    kc = L["channels"] // H
    out = np.zeros((T, L["channels"]), f32)
    idx = np.arange(T)
    rel = idx[None, :] - idx[:, None]  # j - i
    for h in range(H):
        qs = (q[:, h * kc:(h + 1) * kc] / np.sqrt(f32(kc))).astype(f32)
        kh, vh = k[:, h * kc:(h + 1) * kc], v[:, h * kc:(h + 1) * kc]
        score = qs @ kh.T
        if w > 0:
            rl = qs @ L["embRelK"].T  # [T][2w+1]
            band = np.abs(rel) <= w
            score = score + np.where(band, rl[idx[:, None], np.clip(rel + w, 0, 2 * w)], f32(0))
        p = softmax_rows(score)
        o = p @ vh
        if w > 0:
            pw = np.zeros((T, 2 * w + 1), f32)
            for r in range(2 * w + 1):
                j = idx + r - w
                ok = (j >= 0) & (j < T)
                pw[ok, r] = p[idx[ok], j[ok]]
            o = o + pw @ L["embRelV"]
        out[:, h * kc:(h + 1) * kc] = o
    return conv1d(out, L["o"])


def ffn(x, L):
    """FFN::forward + same_padding, src/modules/ffn.cpp:47-77."""
    k = L["k"]
    pl, pr = (k - 1) // 2, k // 2

    def same(a):
        if k == 1:
            return a
        z = np.zeros((a.shape[0] + pl + pr, a.shape[1]), f32)
        z[pl:pl + a.shape[0]] = a
        return z

    h = relu(conv1d(same(x), L["conv1"]))
    return conv1d(same(h), L["conv2"])


def text_encoder(ids, E):
    """TextEncoder::forward, src/models/TextEncoder.cpp:50-74 and attention_encoder::forward,
    src/modules/attention_encoder.cpp:78-94 (post-LN)."""
    x = (E["emb"][np.asarray(ids)] * np.sqrt(f32(E["hidden"]))).astype(f32)
    for i in range(E["nLayers"]):
        y = mha(x, E["mha"][i])
        x1 = layer_norm(x + y, E["norm1"][i])
        y = ffn(x1, E["ffn"][i])
        x = layer_norm(x1 + y, E["norm2"][i])
    stat = conv1d(x, E["proj"])
    half = stat.shape[1] // 2
    return x, stat[:, :half], stat[:, half:]


# ----------------------------------------------------------------------------------------------
# duration predictors
# ----------------------------------------------------------------------------------------------
def fix_dp(x, D, g=None):
    """FixDurationPredictor::forward, src/models/FixDurationPredictor.cpp:75-96."""
    if g is not None:
        x = x + conv1d(g, D["cond"])
    x = layer_norm(relu(conv1d(x, D["conv_1"])), D["norm_1"])
    x = layer_norm(relu(conv1d(x, D["conv_2"])), D["norm_2"])
    return conv1d(x, D["proj"])


def dds_conv(x, D, g=None):
    """DDSConv::forward, src/modules/DDSConv.cpp:84-111."""
    if g is not None:
        x = x + g
    for i in range(D["nLayers"]):
        y = gelu(layer_norm(conv1d(x, D["sep"][i]), D["norm1"][i]))
        y = gelu(layer_norm(conv1d(y, D["c11"][i]), D["norm2"][i]))
        x = x + y
    return x.astype(f32)


def rq_spline_inverse(x, uw, uh, ud, tail=5.0):
    """unconstrained_rational_quadratic_spline (inverse), src/modules/ConvFlow.cpp:80-240 and
    searchsorted :57-78.  x [T], uw/uh [T][10], ud [T][9]."""
    x = np.asarray(x, f32)
    T, nb = uw.shape
    inside = (x < f32(tail)) & (x > f32(-tail))
    udp = np.full((T, ud.shape[1] + 2), f32(0.5397424172369522), f32)
    udp[:, 1:-1] = ud
    uwm = np.where(inside[:, None], uw, f32(0)).astype(f32)
    uhm = np.where(inside[:, None], uh, f32(0)).astype(f32)
    mbw = mbh = md = f32(1e-3)
    widths = (softmax_rows(uwm) * (f32(1) - mbw * nb) + mbw).astype(f32)
    cw = np.zeros((T, nb + 1), f32)
    cw[:, 1:] = np.cumsum(widths, axis=1, dtype=f32)
    cw = (cw * f32(2 * tail) + f32(-tail)).astype(f32)
    cw[:, 0], cw[:, -1] = -tail, tail
    wsub = cw[:, 1:] - cw[:, :-1]
    dv = (softplus(udp) + md).astype(f32)
    heights = (softmax_rows(uhm) * (f32(1) - mbh * nb) + mbh).astype(f32)
    ch = np.zeros((T, nb + 1), f32)
    ch[:, 1:] = np.cumsum(heights, axis=1, dtype=f32)
    ch = (ch * f32(2 * tail) + f32(-tail)).astype(f32)
    ch[:, 0], ch[:, -1] = -tail, tail
    hts = ch[:, 1:] - ch[:, :-1]
    loc = ch.copy()
    loc[:, -1] += f32(1e-6)
    bidx = (x[:, None] >= loc).sum(axis=1) - 1
    bidx = np.clip(bidx, 0, nb - 1)  # rows outside the tails are overwritten below
    r = np.arange(T)
    delta = hts / wsub
    icw, ibw, ich = cw[r, bidx], wsub[r, bidx], ch[r, bidx]
    idl, idv, idv1, ih = delta[r, bidx], dv[r, bidx], dv[r, bidx + 1], hts[r, bidx]
    with np.errstate(all="ignore"):
        a = (x - ich) * (idv + idv1 - idl * 2) + ih * (idl - idv)
        b = ih * idv - (x - ich) * (idv + idv1 - 2 * idl)
        c = -(idl * (x - ich))
        disc = b * b - a * c * 4
        root = (c * 2) / (-b - np.sqrt(disc))
        out = root * ibw + icw
    return np.where(inside, out, x).astype(f32)


def conv_flow(x2, XX, CF):
    """ConvFlow::forward, src/modules/ConvFlow.cpp:242-270. x2: [2][T] -> returns [T][2]."""
    x0, x1 = x2[0][:, None].astype(f32), x2[1].astype(f32)
    h = conv1d(x0, CF["pre"])
    h = dds_conv(h, CF["dds"], g=XX)
    h = conv1d(h, CF["proj"])
    fs = np.sqrt(f32(CF["pre"]["outCh"]))
    uw, uh, ud = (h[:, :10] / fs).astype(f32), (h[:, 10:20] / fs).astype(f32), h[:, 20:]
    y1 = rq_spline_inverse(x1, uw, uh, ud)
    return np.stack([x0[:, 0], y1], axis=1).astype(f32)


def stoch_dp(x, D, g=None):
    """StochasticDurationPredictor::forward, src/models/StochasticDurationPredictor.cpp:117-149
    with noiseScale == 0 (z == 0); flows[0] is skipped (:138)."""
    XX = conv1d(x, D["pre"])
    if g is not None:
        XX = XX + conv1d(g, D["cond"])
    XX = dds_conv(XX, D["convs"])
    XX = conv1d(XX, D["proj"])
    T = XX.shape[0]
    zT = np.zeros((2, T), f32)  # nn_flip(z, 0) of zeros
    flap = zT.T
    for i in range(D["nFlows"] - 1, 0, -1):
        flap = conv_flow(zT, XX, D["flows"][i])  # [T][2]
        flap = flap[:, ::-1]                      # nn_flip(.,1): reverse channels
        zT = flap.T.copy()
    # ElementwiseAffine::forward, src/modules/ElementwiseAffine.cpp:44-58
    aff = (flap - D["ea"]["m"][None, :]) * np.exp(-D["ea"]["logs"])[None, :]
    return aff[:, :1].astype(f32)


# ----------------------------------------------------------------------------------------------
# flow
# ----------------------------------------------------------------------------------------------
def wn(x, Wn, g=None):
    """WN::forward + fused_add_tanh_sigmoid_multiply, src/modules/WN.cpp:85-149."""
    H = Wn["in_layers"][0]["inCh"]
    out = np.zeros_like(x)
    gg = conv1d(g, Wn["cond"]) if g is not None else None
    n = Wn["nLayers"]
    for i in range(n):
        a = conv1d(x, Wn["in_layers"][i])
        if gg is not None:
            a = a + gg[:, i * 2 * H:(i + 1) * 2 * H]
        acts = tanh_ref(a[:, :H]) * sigmoid(a[:, H:])
        rs = conv1d(acts, Wn["res_skip"][i])
        if i < n - 1:
            x = x + rs[:, :H]
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out.astype(f32)


def flow_reverse(z, Fl, g=None):
    """ResidualCouplingBlock::forward, src/models/ResidualCouplingBlock.cpp:59-71 and
    ResidualCouplingLayer::forward, src/modules/ResidualCouplingLayer.cpp:47-66."""
    x = np.asarray(z, f32)
    for i in range(Fl["nFlows"] - 1, -1, -1):
        x = x[:, ::-1]
        L = Fl["layers"][i]
        half = x.shape[1] // 2
        x0, x1 = x[:, :half], x[:, half:]
        h = wn(conv1d(x0, L["pre"]), L["wn"], g)
        m = conv1d(h, L["post"])
        x = np.concatenate([x0, x1 - m], axis=1).astype(f32)
    return x


# ----------------------------------------------------------------------------------------------
# decoders
# ----------------------------------------------------------------------------------------------
HANN = np.array([0.0, 0.03806023, 0.14644661, 0.30865828, 0.5, 0.69134172, 0.85355339, 0.96193977, 1.0,
                 0.96193977, 0.85355339, 0.69134172, 0.5, 0.30865828, 0.14644661, 0.03806023], f32)  # hann.cpp:3-5
HANN_POW = np.array([0.0, 0.00144858, 0.02144661, 0.09526994, 0.25, 0.47795337, 0.72855339, 0.92532811, 1.0,
                     0.92532811, 0.72855339, 0.47795337, 0.25, 0.09526994, 0.02144661, 0.00144858], f32)  # :6-9

PQMF_PROTO = np.array([
    8.36595339e-06, 2.68017852e-05, 5.05711124e-05, 6.13482515e-05, 2.75281598e-05, -8.62839965e-05,
    -2.99268467e-04, -5.88389492e-04, -8.67064627e-04, -9.82905838e-04, -7.47200209e-04, 8.04087656e-19,
    1.30001234e-03, 2.98798828e-03, 4.64603942e-03, 5.63488600e-03, 5.22586317e-03, 2.82493436e-03,
    -1.75650987e-03, -8.06073440e-03, -1.48622207e-02, -2.02404650e-02, -2.18780344e-02, -1.75512321e-02,
    -5.71474631e-03, 1.39652689e-02, 4.02848855e-02, 7.05021626e-02, 1.00706377e-01, 1.26503321e-01,
    1.43873012e-01, 1.50000000e-01, 1.43873012e-01, 1.26503321e-01, 1.00706377e-01, 7.05021626e-02,
    4.02848855e-02, 1.39652689e-02, -5.71474631e-03, -1.75512321e-02, -2.18780344e-02, -2.02404650e-02,
    -1.48622207e-02, -8.06073440e-03, -1.75650987e-03, 2.82493436e-03, 5.22586317e-03, 5.63488600e-03,
    4.64603942e-03, 2.98798828e-03, 1.30001234e-03, 8.04087656e-19, -7.47200209e-04, -9.82905838e-04,
    -8.67064627e-04, -5.88389492e-04, -2.99268467e-04, -8.62839965e-05, 2.75281598e-05, 6.13482515e-05,
    5.05711124e-05, 2.68017852e-05, 8.36595339e-06], f32)  # pqmf.cpp:8-25


def pqmf_synthesis_filters():
    """pqmf::pqmf, src/modules/pqmf.cpp:39-95: h_k[n] = 2 h[n] cos((2k+1) pi/8 (n - 30.5) - (-1)^k pi/4),
    evaluated in float32 like the reference (Eigen float arrays)."""
    n = np.arange(63, dtype=f32)
    tmp1 = ((n - f32(61.0 / 2.0)) * f32(np.pi / 8.0)).astype(f32)
    H = np.zeros((4, 63), f32)
    for k in range(4):
        ph = f32(((-1.0) ** k) * (np.pi / 4))
        H[k] = PQMF_PROTO * f32(2) * np.cos((tmp1 * f32(2 * k + 1) - ph).astype(f32)).astype(f32)
    return H


def istft(mag, phase):
    """iStft::forward (16,4,16), src/modules/iStft.cpp:46-124; the real inverse FFT uses only Re of
    the DC and Nyquist bins (eigen unsupported/Eigen/src/FFT/ei_kissfft_impl.h:376-405), scale 1/16."""
    mag, phase = np.asarray(mag, f32), np.asarray(phase, f32)
    Fr = mag.shape[0]
    re, im = mag * np.cos(phase), mag * np.sin(phase)
    n = np.arange(16)
    k = np.arange(1, 8)
    ang = 2 * np.pi * np.outer(k, n) / 16.0
    x = (re[:, :1] + re[:, 8:9] * ((-1.0) ** n)[None, :] + 2 * (re[:, 1:8] @ np.cos(ang) - im[:, 1:8] @ np.sin(ang))) / 16.0
    x = (x.astype(f32) * HANN[None, :]).astype(f32)
    L = (Fr - 1) * 4 + 16
    ret, ws = np.zeros(L, f32), np.zeros(L, f32)
    for j in range(Fr):
        ret[4 * j:4 * j + 16] += x[j]
        ws[4 * j:4 * j + 16] += HANN_POW
    ret = np.where(ws > 1e-14, ret / np.where(ws > 1e-14, ws, 1), ret).astype(f32)
    return ret[8:8 + (Fr - 1) * 4]


def resblock1(x, R):
    """ResBlock1::forward, src/modules/ResBlock1.cpp:55-69."""
    for i in range(R["n"]):
        xt = conv1d(leaky_relu(x, 0.1), R["convs1"][i])
        xt = conv1d(leaky_relu(xt, 0.1), R["convs2"][i])
        x = xt + x
    return x.astype(f32)


def _gen_trunk(z, G, g=None):
    """Shared conv_pre -> [leaky(0.1) -> ConvTranspose -> mean of ResBlock1]xN -> leaky(0.01) trunk of
    Generator_hifigan.cpp:139-178 / Generator_MS.cpp:166-198 / Generator_Istft.cpp:149-180 /
    Generator_MBB.cpp:143-175."""
    x = conv1d(z, G["conv_pre"])
    if g is not None and "cond" in G:
        x = x + conv1d(g, G["cond"])
    nk = len(G["rbK"])
    for i, up in enumerate(G["ups"]):
        x = conv1d_transposed(leaky_relu(x, 0.1), up)
        xs = None
        for j in range(nk):
            r = resblock1(x, G["resblocks"][i * nk + j])
            xs = r if xs is None else xs + r
        x = (xs / f32(nk)).astype(f32)
    return leaky_relu(x, 0.01)


def _subband_frames(x, G):
    """reflect-pad(1,0) + subband_conv_post, Generator_MS.cpp:200-208."""
    xp = np.zeros((x.shape[0] + 1, x.shape[1]), f32)
    xp[1:] = x
    if x.shape[0] > 1:
        xp[0] = x[1]
    return conv1d(xp, G["subband_conv_post"])


def _bands_to_time(s, nb):
    """exp / pi*sin + per-band iSTFT, Generator_MS.cpp:210-223."""
    cols = s.shape[1] // nb
    t = np.zeros(((s.shape[0] - 1) * 4, nb), f32)
    for b in range(nb):
        sb = s[:, b * cols:(b + 1) * cols]
        t[:, b] = istft(np.exp(sb[:, :9]), np.sin(sb[:, 9:18]) * f32(np.pi))
    return t


def _zero_stuff4(t):
    """upDownConv_: ConvTranspose1d stride 4 with weight 4*delta, Generator_MS.cpp:106-124."""
    u = np.zeros((t.shape[0] * 4, t.shape[1]), f32)
    u[::4] = t * f32(4)
    return u


def generator(z, G, g=None):
    dt = G["decType"]
    x = _gen_trunk(z, G, g)
    if dt == 0:  # Generator_hifiGan::forward, Generator_hifigan.cpp:176-181
        return tanh_ref(conv1d(x, G["conv_post"]))[:, 0]
    s = _subband_frames(x, G)
    if dt == 2:  # Generator_Istft::forward, Generator_Istft.cpp:189-197
        return istft(np.exp(s[:, :9]), np.sin(s[:, 9:18]) * f32(np.pi))
    t = _bands_to_time(s, G["subBands"])
    u = _zero_stuff4(t)
    if dt == 1:  # Generator_MS.cpp:225-228
        return conv1d(u, G["multistream_conv_post"])[:, 0]
    # Generator_MBB.cpp:200-202 -> pqmf::forward, pqmf.cpp:97-115
    Hs = pqmf_synthesis_filters()
    cv = dict(outCh=1, inCh=4, k=63, pad=31, dil=1, hasBias=0, w=Hs.T.reshape(1, 63, 4), b=None)
    return conv1d(u, cv)[:, 0]


# ----------------------------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------------------------
def infer(M, ids, sid=0, length_scale=1.0, forced_w=None):
    """NN half of SynthesizerTrn::infer, src/models/SynthesizerTrn.cpp:357-396 (+ expandM :304-321)."""
    ids = np.asarray(ids, np.int64)
    xx, m, _logs = text_encoder(ids, M["enc"])
    g = None
    if M["isMS"] == 1:
        if sid < 0 or sid >= M["spkNum"]:
            sid = 0
        g = M["emg"][sid][None, :].astype(f32)
    logw = (fix_dp if M["durPredType"] == 1 else stoch_dp)(xx, M["dp"], g)
    w = np.exp(logw[:, 0]) * f32(length_scale)
    w_ceil = np.ceil(w).astype(f32)
    if forced_w is not None:
        w_ceil = np.asarray(forced_w, f32)
    reps = w_ceil.astype(np.int64)
    F = max(int(w_ceil.sum()), 1)
    z_p = np.zeros((F, m.shape[1]), f32)
    rows = np.repeat(np.arange(len(ids)), reps)
    z_p[:len(rows)] = m[rows]
    z = flow_reverse(z_p, M["flow"], g)
    o = generator(z, M["dec"], g)
    pcm = np.trunc(o * f32(32737)).astype(np.int16)  # (int16_t)(o*32737), SynthesizerTrn.cpp:395
    return dict(xx=xx, m=m, logw=logw[:, 0], w_ceil=w_ceil, z_p=z_p, z=z, o=o, pcm=pcm, F=F)
