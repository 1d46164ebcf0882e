// model.hpp — host-side parser of the SummerTTS `.bin` NN section (pure C++, no CUDA).
//
// Walks the flat float32 stream in exactly the order the reference constructors consume it
// (reference: src/models/SynthesizerTrn.cpp:101-167 and the sub-constructors
// cited at each record below) and produces a typed layer tree of views into the caller's blob.
// Nothing is copied here; the engine repacks the weights into device layouts afterwards.
#pragma once
#include <cmath>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace stts {

struct FormatError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Cursor {
    const float* p;
    int64_t n, o = 0;
    int32_t i() {      // integer stored as float: must be finite and representable (a cast of NaN / inf / 1e30 is undefined)
        need(1);
        const float v = p[o++];
        if (!(v > -2147483648.0f && v < 2147483648.0f)) throw FormatError("non-integer header value at float offset " + std::to_string(o - 1));
        return (int32_t)v;
    }
    const float* f(int64_t cnt) {
        need(cnt);
        const float* r = p + o;
        o += cnt;
        return r;
    }
    void need(int64_t cnt) const {
        if (cnt < 0 || o + cnt > n) throw FormatError("model blob truncated at float offset " + std::to_string(o));
    }
};

// conv1d := outCh inCh k pad dil hasBias W[outCh][k][inCh] [b[outCh]]     nn_conv1d.cpp:25-52
struct ConvRec {
    int32_t outCh = 0, inCh = 0, k = 0, pad = 0, dil = 1, hasBias = 0, stride = 1, sep = 0;
    const float* w = nullptr;  // file order [o][k][c]
    const float* b = nullptr;
};
// lnorm := size gamma beta                                                 nn_layer_norm.cpp:18-34
struct LNormRec {
    int32_t size = 0;
    const float *gamma = nullptr, *beta = nullptr;
};
// mha := channels outCh nHeads winSize [r c Ek] [r c Ev] q k v o           multi_head_attention.cpp:40-90
struct MhaRec {
    int32_t channels = 0, outCh = 0, nHeads = 0, winSize = 0, relRows = 0, relCols = 0;
    const float *embRelK = nullptr, *embRelV = nullptr;  // col-major (rows x cols): e(r,c) = p[c*rows + r]
    ConvRec q, k, v, o;
};
struct FfnRec {  // ffn := k conv1d conv1d                                   ffn.cpp:27-30
    int32_t k = 0;
    ConvRec c1, c2;
};
struct EncRec {  // TextEncoder.cpp:32-44, attention_encoder.cpp:30-55
    int32_t hidden = 0, vocab = 0, embDim = 0, nLayers = 0;
    const float* emb = nullptr;  // col-major (vocab x embDim): e(v,c) = p[c*vocab + v]
    std::vector<MhaRec> mha;
    std::vector<LNormRec> norm1, norm2;
    std::vector<FfnRec> ffn;
    ConvRec proj;
};
struct ResBlockRec {  // ResBlock1.cpp:27-38
    int32_t n = 0;
    std::vector<ConvRec> convs1, convs2;
};
struct GenRec {  // Generator_MS.cpp:51-127, Generator_hifigan.cpp:44-101, Generator_Istft.cpp, Generator_MBB.cpp
    int32_t decType = 0, subBands = 0, nfft = 0, hop = 0, upInitCh = 0;
    std::vector<int32_t> upRates, upK, rbK;
    std::vector<std::vector<int32_t>> rbD;
    ConvRec conv_pre, conv_post, cond, subband_post, ms_post;
    bool hasCond = false;
    std::vector<ConvRec> ups;
    std::vector<ResBlockRec> rbs;
};
struct WnRec {  // WN.cpp:32-60
    int32_t nLayers = 0, k = 0;
    std::vector<ConvRec> in_layers, res_skip;
    ConvRec cond;
    bool hasCond = false;
};
struct CouplingRec {  // ResidualCouplingLayer.cpp:28-30
    ConvRec pre, post;
    WnRec wn;
};
struct FlowRec {  // ResidualCouplingBlock.cpp:29-39
    int32_t nFlows = 0, nLayers = 0;
    std::vector<CouplingRec> layers;
};
struct DdsRec {  // DDSConv.cpp:29-59
    int32_t nLayers = 0, k = 0;
    std::vector<ConvRec> sep, c11;
    std::vector<LNormRec> norm1, norm2;
};
struct EaRec {  // ElementwiseAffine.cpp:27-33
    const float *m = nullptr, *logs = nullptr;
};
struct ConvFlowRec {  // ConvFlow.cpp:38-40
    ConvRec pre, proj;
    DdsRec dds;
};
struct DurPredRec {  // FixDurationPredictor.cpp:33-44 / StochasticDurationPredictor.cpp:41-70
    int32_t type = 1;
    // fix
    ConvRec conv_1, conv_2, proj, cond;
    LNormRec norm_1, norm_2;
    bool hasCond = false;
    // stochastic
    int32_t nFlows = 0;
    EaRec ea, post_ea;
    std::vector<ConvFlowRec> flows, post_flows;
    ConvRec post_pre, post_proj, pre;
    DdsRec post_convs, convs;
};
struct Model {
    int32_t isMS = 0, langType = 0, durPredType = 0, decType = 0, spkNum = 0, gin = 0;
    EncRec enc;
    GenRec dec;
    FlowRec flow;
    DurPredRec dp;
    const float* emg = nullptr;  // col-major (spkNum x gin): e(s,c) = p[c*spkNum + s]
    int64_t nnEnd = 0;
};

inline ConvRec parse_conv(Cursor& c) {
    ConvRec r;
    r.outCh = c.i(); r.inCh = c.i(); r.k = c.i(); r.pad = c.i(); r.dil = c.i(); r.hasBias = c.i();
    if (r.outCh <= 0 || r.inCh <= 0 || r.k <= 0) throw FormatError("bad conv1d record at " + std::to_string(c.o));
    r.w = c.f((int64_t)r.outCh * r.k * r.inCh);
    if (r.hasBias == 1) r.b = c.f(r.outCh);
    return r;
}
inline ConvRec parse_convT(Cursor& c) {  // nn_conv1d_transposed.cpp:25-52
    ConvRec r;
    r.outCh = c.i(); r.inCh = c.i(); r.k = c.i(); r.pad = c.i(); r.dil = c.i(); r.hasBias = c.i(); r.stride = c.i();
    if (r.outCh <= 0 || r.inCh <= 0 || r.k <= 0) throw FormatError("bad convT record at " + std::to_string(c.o));
    r.w = c.f((int64_t)r.outCh * r.k * r.inCh);
    if (r.hasBias == 1) r.b = c.f(r.outCh);
    return r;
}
inline LNormRec parse_ln(Cursor& c) {
    LNormRec r;
    r.size = c.i();
    r.gamma = c.f(r.size);
    r.beta = c.f(r.size);
    return r;
}
inline MhaRec parse_mha(Cursor& c) {
    MhaRec m;
    m.channels = c.i(); m.outCh = c.i(); m.nHeads = c.i(); m.winSize = c.i();
    if (m.winSize != 0) {
        m.relRows = c.i(); m.relCols = c.i();
        m.embRelK = c.f((int64_t)m.relRows * m.relCols);
        int32_t r2 = c.i(), c2 = c.i();
        if (r2 != m.relRows || c2 != m.relCols) throw FormatError("rel-k / rel-v shape mismatch");
        m.embRelV = c.f((int64_t)r2 * c2);
    }
    m.q = parse_conv(c); m.k = parse_conv(c); m.v = parse_conv(c); m.o = parse_conv(c);
    return m;
}
inline EncRec parse_enc(Cursor& c) {
    EncRec e;
    e.hidden = c.i(); e.vocab = c.i(); e.embDim = c.i();
    e.emb = c.f((int64_t)e.vocab * e.embDim);
    e.nLayers = c.i();
    if (e.nLayers < 0 || e.nLayers > 64) throw FormatError("bad encoder layer count");
    for (int i = 0; i < e.nLayers; ++i) e.mha.push_back(parse_mha(c));
    for (int i = 0; i < e.nLayers; ++i) e.norm1.push_back(parse_ln(c));
    for (int i = 0; i < e.nLayers; ++i) {
        FfnRec f;
        f.k = c.i(); f.c1 = parse_conv(c); f.c2 = parse_conv(c);
        e.ffn.push_back(f);
    }
    for (int i = 0; i < e.nLayers; ++i) e.norm2.push_back(parse_ln(c));
    e.proj = parse_conv(c);
    return e;
}
inline ResBlockRec parse_rb(Cursor& c) {
    ResBlockRec r;
    r.n = c.i();
    if (r.n < 0 || r.n > 16) throw FormatError("bad resblock size");
    for (int i = 0; i < r.n; ++i) r.convs1.push_back(parse_conv(c));
    for (int i = 0; i < r.n; ++i) r.convs2.push_back(parse_conv(c));
    return r;
}
inline GenRec parse_gen(Cursor& c, int decType, int isMS) {
    GenRec g;
    g.decType = decType;
    if (decType >= 1) { g.subBands = c.i(); g.nfft = c.i(); g.hop = c.i(); }
    int nUp = c.i();
    if (nUp < 0 || nUp > 16) throw FormatError("bad upsample count");
    for (int i = 0; i < nUp; ++i) g.upRates.push_back(c.i());
    g.upInitCh = c.i();
    int nUpK = c.i();
    if (nUpK < 0 || nUpK > 64) throw FormatError("implausible upsample kernel count");
    for (int i = 0; i < nUpK; ++i) g.upK.push_back(c.i());
    int nRbK = c.i();
    if (nRbK < 1 || nRbK > 16) throw FormatError("implausible resblock kernel count");
    for (int i = 0; i < nRbK; ++i) g.rbK.push_back(c.i());
    int nRbD = c.i();
    if (nRbD < 0 || nRbD > 64 || (nRbD != 0 && nRbD != nRbK)) throw FormatError("resblock dilation list does not match the kernel list");
    for (int i = 0; i < nRbD; ++i) { int a = c.i(), b = c.i(), d = c.i(); g.rbD.push_back({a, b, d}); }
    if (nUpK < nUp) throw FormatError("fewer upsample kernel sizes than rates");
    g.conv_pre = parse_conv(c);
    for (int i = 0; i < nUp; ++i) {
        ConvRec u = parse_convT(c);
        u.stride = g.upRates[i];                                   // ctor override, Generator_MS.cpp:88-93
        u.pad = (int)std::floor((float)(g.upK[i] - g.upRates[i]) / 2.0f);
        g.ups.push_back(u);
    }
    for (int i = 0; i < nUp * nRbK; ++i) g.rbs.push_back(parse_rb(c));
    if (decType == 0) {
        g.conv_post = parse_conv(c);
        if (isMS == 1) { g.cond = parse_conv(c); g.hasCond = true; }
    } else {
        g.subband_post = parse_conv(c);
        if (decType == 1) g.ms_post = parse_conv(c);
    }
    return g;
}
inline WnRec parse_wn(Cursor& c, int isMS) {
    WnRec w;
    w.nLayers = c.i(); w.k = c.i();
    if (w.nLayers < 0 || w.nLayers > 64) throw FormatError("bad WN layer count");
    for (int i = 0; i < w.nLayers; ++i) {
        ConvRec cv = parse_conv(c);
        cv.dil = 1;                        // WN.cpp:36-42, dilation_rate == 1 (SynthesizerTrn.cpp:134)
        cv.pad = (w.k * 1 - 1) / 2;
        w.in_layers.push_back(cv);
    }
    for (int i = 0; i < w.nLayers; ++i) w.res_skip.push_back(parse_conv(c));
    if (isMS == 1) { w.cond = parse_conv(c); w.hasCond = true; }
    return w;
}
inline FlowRec parse_flow(Cursor& c, int isMS) {
    FlowRec f;
    f.nFlows = c.i(); f.nLayers = c.i();
    if (f.nFlows < 0 || f.nFlows > 64) throw FormatError("bad flow count");
    for (int i = 0; i < f.nFlows; ++i) {
        CouplingRec L;
        L.pre = parse_conv(c);
        L.wn = parse_wn(c, isMS);
        L.post = parse_conv(c);
        f.layers.push_back(L);
    }
    return f;
}
inline DdsRec parse_dds(Cursor& c) {
    DdsRec d;
    d.nLayers = c.i(); d.k = c.i();
    if (d.nLayers < 0 || d.nLayers > 16) throw FormatError("bad DDSConv layer count");
    int dil = 1;
    for (int i = 0; i < d.nLayers; ++i) {
        ConvRec cv = parse_conv(c);
        cv.dil = dil;                                                     // DDSConv.cpp:33-41
        cv.pad = (int)std::floor((float)(d.k * dil - dil) / 2.0f);
        cv.sep = 1;
        d.sep.push_back(cv);
        dil *= d.k;
    }
    for (int i = 0; i < d.nLayers; ++i) d.c11.push_back(parse_conv(c));
    for (int i = 0; i < d.nLayers; ++i) d.norm1.push_back(parse_ln(c));
    for (int i = 0; i < d.nLayers; ++i) d.norm2.push_back(parse_ln(c));
    return d;
}
inline EaRec parse_ea(Cursor& c) {
    EaRec e;
    e.m = c.f(2);
    e.logs = c.f(2);
    return e;
}
inline ConvFlowRec parse_convflow(Cursor& c) {
    ConvFlowRec f;
    f.pre = parse_conv(c);
    f.dds = parse_dds(c);
    f.proj = parse_conv(c);
    return f;
}
inline DurPredRec parse_dp(Cursor& c, int type, int isMS) {
    DurPredRec d;
    d.type = type;
    if (type == 1) {
        d.conv_1 = parse_conv(c); d.norm_1 = parse_ln(c);
        d.conv_2 = parse_conv(c); d.norm_2 = parse_ln(c);
        d.proj = parse_conv(c);
        if (isMS == 1) { d.cond = parse_conv(c); d.hasCond = true; }
        return d;
    }
    d.nFlows = c.i();
    if (d.nFlows < 0 || d.nFlows > 16) throw FormatError("bad SDP flow count");
    d.ea = parse_ea(c);
    for (int i = 0; i < d.nFlows; ++i) d.flows.push_back(parse_convflow(c));
    d.post_pre = parse_conv(c);
    d.post_proj = parse_conv(c);
    d.post_convs = parse_dds(c);
    d.post_ea = parse_ea(c);
    for (int i = 0; i < 4; ++i) d.post_flows.push_back(parse_convflow(c));
    d.pre = parse_conv(c);
    d.proj = parse_conv(c);
    d.convs = parse_dds(c);
    if (isMS == 1) { d.cond = parse_conv(c); d.hasCond = true; }
    return d;
}

inline Model parse_model(const float* blob, int64_t nfloats) {
    Cursor c{blob, nfloats};
    Model M;
    M.isMS = c.i(); M.langType = c.i(); M.durPredType = c.i(); M.decType = c.i();
    if (M.decType < 0 || M.decType > 3) throw FormatError("unknown decoder type " + std::to_string(M.decType));
    if (M.durPredType < 0 || M.durPredType > 1)
        throw FormatError("unknown duration predictor type " + std::to_string(M.durPredType));
    M.enc = parse_enc(c);
    M.dec = parse_gen(c, M.decType, M.isMS);
    M.flow = parse_flow(c, M.isMS);
    M.dp = parse_dp(c, M.durPredType, M.isMS);
    if (M.isMS == 1) {
        M.spkNum = c.i(); M.gin = c.i();
        if (M.spkNum <= 0 || M.gin <= 0) throw FormatError("bad speaker table header");
        M.emg = c.f((int64_t)M.spkNum * M.gin);
    }
    M.nnEnd = c.o;
    return M;
}

// ---- text description (used by the CPU tests to cross-check against binfmt.py) -------------
inline void desc_conv(std::ostringstream& s, const char* name, const ConvRec& c) {
    s << name << " conv " << c.inCh << "->" << c.outCh << " k" << c.k << " p" << c.pad << " d" << c.dil << " b"
      << c.hasBias;
    if (c.stride != 1) s << " s" << c.stride;
    if (c.sep) s << " sep";
    s << "\n";
}
inline std::string describe(const Model& M) {
    std::ostringstream s;
    s << "header isMS=" << M.isMS << " lang=" << M.langType << " dp=" << M.durPredType << " dec=" << M.decType
      << " spk=" << M.spkNum << " gin=" << M.gin << " nn_end=" << M.nnEnd << "\n";
    s << "enc hidden=" << M.enc.hidden << " vocab=" << M.enc.vocab << " emb=" << M.enc.embDim
      << " layers=" << M.enc.nLayers << "\n";
    for (auto& m : M.enc.mha) {
        s << "mha ch=" << m.channels << " heads=" << m.nHeads << " win=" << m.winSize << "\n";
        desc_conv(s, "q", m.q); desc_conv(s, "k", m.k); desc_conv(s, "v", m.v); desc_conv(s, "o", m.o);
    }
    for (auto& f : M.enc.ffn) { desc_conv(s, "ffn1", f.c1); desc_conv(s, "ffn2", f.c2); }
    desc_conv(s, "enc_proj", M.enc.proj);
    desc_conv(s, "conv_pre", M.dec.conv_pre);
    for (auto& u : M.dec.ups) desc_conv(s, "up", u);
    for (auto& r : M.dec.rbs) {
        for (auto& cv : r.convs1) desc_conv(s, "rb1", cv);
        for (auto& cv : r.convs2) desc_conv(s, "rb2", cv);
    }
    if (M.dec.decType == 0) desc_conv(s, "conv_post", M.dec.conv_post);
    else desc_conv(s, "subband_post", M.dec.subband_post);
    if (M.dec.decType == 1) desc_conv(s, "ms_post", M.dec.ms_post);
    if (M.dec.hasCond) desc_conv(s, "dec_cond", M.dec.cond);
    for (auto& L : M.flow.layers) {
        desc_conv(s, "flow_pre", L.pre);
        for (auto& cv : L.wn.in_layers) desc_conv(s, "wn_in", cv);
        for (auto& cv : L.wn.res_skip) desc_conv(s, "wn_rs", cv);
        if (L.wn.hasCond) desc_conv(s, "wn_cond", L.wn.cond);
        desc_conv(s, "flow_post", L.post);
    }
    if (M.dp.type == 1) {
        desc_conv(s, "dp_conv1", M.dp.conv_1); desc_conv(s, "dp_conv2", M.dp.conv_2); desc_conv(s, "dp_proj", M.dp.proj);
    } else {
        s << "sdp flows=" << M.dp.nFlows << "\n";
        desc_conv(s, "sdp_pre", M.dp.pre); desc_conv(s, "sdp_proj", M.dp.proj);
        for (auto& cv : M.dp.convs.sep) desc_conv(s, "sdp_sep", cv);
        for (auto& f : M.dp.flows) { desc_conv(s, "cf_pre", f.pre); desc_conv(s, "cf_proj", f.proj); }
    }
    if (M.dp.hasCond) desc_conv(s, "dp_cond", M.dp.cond);
    return s.str();
}

}  // namespace stts
