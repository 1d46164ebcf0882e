// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// ID-level driver around the UNMODIFIED reference objects (compiled in place from
// /root/reference by oracle/Makefile).  It restates
//   * the NN half of the constructor  /root/reference/src/models/SynthesizerTrn.cpp:101-167
//   * the NN half of infer()          /root/reference/src/models/SynthesizerTrn.cpp:357-396
// over the reference's own TextEncoder / *DurationPredictor / ResidualCouplingBlock /
// Generator_* classes (public headers in /root/reference/src/header), so that phoneme-ID
// sequences can be fed directly (the reference's infer() only accepts text) and every
// stage tensor can be dumped.  It also exposes op-level entry points (conv1d, conv-transpose,
// layer-norm, iSTFT, PQMF ...) that construct the reference's nn_op/module classes from
// caller-provided weights, for known-answer tests of single kernels.
//
// Exposed as a C ABI (libstts_ref.so) consumed by tests/, bench.py's cpu_baseline /
// --impl reference arm and __graft_entry__.smoke() via ctypes.
//
// All dumped matrices are written TIME-MAJOR / channels-last: out[t*C + c].

#include <Eigen/Dense>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "TextEncoder.h"
#include "DurationPredictor_base.h"
#include "FixDurationPredictor.h"
#include "StochasticDurationPredictor.h"
#include "ResidualCouplingBlock.h"
#include "Generator_base.h"
#include "Generator_hifigan.h"
#include "Generator_MS.h"
#include "Generator_Istft.h"
#include "Generator_MBB.h"
#include "nn_conv1d.h"
#include "nn_conv1d_transposed.h"
#include "nn_layer_norm.h"
#include "nn_clamp_min.h"
#include "nn_gelu.h"
#include "nn_tanh.h"
#include "nn_softmax.h"
#include "WN.h"
#include "ResBlock1.h"
#include "iStft.h"
#include "pqmf.h"
#include "ffn.h"
#include "multi_head_attention.h"
#include "DDSConv.h"
#include "ConvFlow.h"

using Eigen::MatrixXf;
using Eigen::Map;

namespace {

struct RefModel {
    int32_t isMS = 0, langType = 0, durPredType = 0, decType = 0;
    int32_t spkNum = 0, gin = 0;
    int64_t nnEnd = 0;
    TextEncoder* enc = nullptr;
    DurationPredictor_base* dp = nullptr;
    ResidualCouplingBlock* flow = nullptr;
    Generator_base* dec = nullptr;
    MatrixXf emg;
};

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// Eigen MatrixXf (rows=time, cols=channels, col-major) -> malloc'd time-major buffer.
float* dump_tm(const MatrixXf& m) {
    float* p = (float*)malloc(sizeof(float) * (size_t)std::max<int64_t>(1, m.rows() * m.cols()));
    for (int64_t t = 0; t < m.rows(); ++t)
        for (int64_t c = 0; c < m.cols(); ++c) p[t * m.cols() + c] = m(t, c);
    return p;
}

MatrixXf load_tm(const float* x, int T, int C) {
    MatrixXf m(T, C);
    for (int t = 0; t < T; ++t)
        for (int c = 0; c < C; ++c) m(t, c) = x[(size_t)t * C + c];
    return m;
}

// same as expandM, SynthesizerTrn.cpp:304-321 (file-static there, so restated)
MatrixXf expand_rows(const MatrixXf& x, const MatrixXf& lengthM) {
    MatrixXf y_lengths = nn_clamp_min(lengthM.colwise().sum(), 1.0);
    int32_t totalLen = (int32_t)y_lengths(0, 0);
    MatrixXf ret = MatrixXf::Zero(totalLen, x.cols());
    int32_t rowIdx = 0;
    for (int32_t i = 0; i < lengthM.rows(); i++) {
        int32_t len = (int32_t)lengthM(i, 0);
        for (int32_t j = 0; j < len; j++) ret.row(rowIdx++) = x.row(i);
    }
    return ret;
}

}  // namespace

extern "C" {

// Stage dump of one utterance.  Every pointer is malloc'd (free with sref_free_result).
typedef struct {
    int32_t T, F, S, hidden, inter;       // ids, frames, samples, encoder width, flow width
    float* xx;      // [T][hidden]   encoder output
    float* m;       // [T][inter]    prior mean
    float* logw;    // [T]
    float* w_ceil;  // [T]
    float* z_p;     // [F][inter]
    float* z;       // [F][inter]
    float* o;       // [S]           raw float waveform
    int16_t* pcm;   // [S]
    double ms[6];   // enc, dp, expand, flow, dec, total
} sref_result;

void sref_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
    Eigen::setNbThreads(n);
#else
    (void)n;
#endif
}

int sref_get_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// SynthesizerTrn.cpp:101-167 (NN section only; the frontend tail is not touched)
void* sref_create(const float* blob, int64_t nfloats) {
    (void)nfloats;
    float* modelData = const_cast<float*>(blob);
    RefModel* M = new RefModel();
    int32_t offset = 0;
    M->isMS = (int32_t)modelData[offset++];
    M->langType = (int32_t)modelData[offset++];
    M->durPredType = (int32_t)modelData[offset++];
    M->decType = (int32_t)modelData[offset++];
    M->enc = new TextEncoder(modelData, offset);
    if (M->decType == 0) M->dec = new Generator_hifiGan(modelData, offset, M->isMS);
    else if (M->decType == 1) M->dec = new Generator_MS(modelData, offset, M->isMS);
    else if (M->decType == 2) M->dec = new Generator_Istft(modelData, offset, M->isMS);
    else if (M->decType == 3) M->dec = new Generator_MBB(modelData, offset, M->isMS);
    else { delete M; return nullptr; }
    M->flow = new ResidualCouplingBlock(modelData, offset, 1, M->isMS);
    if (M->durPredType == 0) M->dp = new StochasticDurationPredictor(modelData, offset, M->isMS);
    else if (M->durPredType == 1) M->dp = new FixDurationPredictor(modelData, offset, M->isMS);
    else { delete M; return nullptr; }
    if (M->isMS == 1) {
        M->spkNum = (int32_t)modelData[offset++];
        M->gin = (int32_t)modelData[offset++];
        M->emg = Map<MatrixXf>(modelData + offset, M->spkNum, M->gin);
        M->dp->setMSSpk(M->isMS, M->gin);
        offset += M->spkNum * M->gin;
    } else {
        M->dp->setMSSpk(0, 0);
    }
    M->nnEnd = offset;
    return M;
}

int64_t sref_nn_end(void* h) { return ((RefModel*)h)->nnEnd; }
int32_t sref_speaker_num(void* h) { return ((RefModel*)h)->spkNum; }
int32_t sref_lang_type(void* h) { return ((RefModel*)h)->langType; }

// SynthesizerTrn.cpp:357-396 with ids given directly.
// forced_w (nullable): per-id frame counts that replace ceil(exp(logw)*lengthScale).
// want_dumps == 0 skips the copies of intermediate tensors (timing runs).
int sref_infer(void* h, const int32_t* ids, int32_t n, int32_t sid, float lengthScale,
               const float* forced_w, int want_dumps, sref_result* R) {
    RefModel* M = (RefModel*)h;
    memset(R, 0, sizeof(*R));
    std::vector<int32_t> idv(ids, ids + n);
    float noiseScale = 0.0;
    double t0 = now_ms();
    MatrixXf m, logs;
    MatrixXf XX = M->enc->forward(idv.data(), n, m, logs);
    MatrixXf g;
    if (M->isMS == 1) {
        if ((sid < 0) || (sid >= M->spkNum)) sid = 0;
        g = M->emg.row(sid);
    }
    double t1 = now_ms();
    MatrixXf logw = M->dp->forward(XX, g, noiseScale);
    double t2 = now_ms();
    MatrixXf w = logw.array().exp() * lengthScale;
    MatrixXf w_ceil = w.array().ceil();
    if (forced_w)
        for (int32_t i = 0; i < n; ++i) w_ceil(i, 0) = forced_w[i];
    MatrixXf m_expand = expand_rows(m, w_ceil);
    // noiseScale == 0: z_p = m_expand (+ randn*logs*0), SynthesizerTrn.cpp:383
    MatrixXf z_p = m_expand;
    double t3 = now_ms();
    MatrixXf z = M->flow->forward(z_p, g);
    double t4 = now_ms();
    MatrixXf o = M->dec->forward(z, g);
    double t5 = now_ms();
    int32_t dataLen = (int32_t)(o.rows() * o.cols());
    R->T = n; R->F = (int32_t)z_p.rows(); R->S = dataLen;
    R->hidden = (int32_t)XX.cols(); R->inter = (int32_t)m.cols();
    R->pcm = (int16_t*)malloc(sizeof(int16_t) * (size_t)std::max(1, dataLen));
    for (int32_t i = 0; i < dataLen; i++) R->pcm[i] = (int16_t)(o.data()[i] * 32737);
    double t6 = now_ms();
    R->ms[0] = t1 - t0; R->ms[1] = t2 - t1; R->ms[2] = t3 - t2; R->ms[3] = t4 - t3;
    R->ms[4] = t5 - t4; R->ms[5] = t6 - t0;
    if (want_dumps) {
        R->xx = dump_tm(XX); R->m = dump_tm(m); R->logw = dump_tm(logw); R->w_ceil = dump_tm(w_ceil);
        R->z_p = dump_tm(z_p); R->z = dump_tm(z);
        R->o = (float*)malloc(sizeof(float) * (size_t)std::max(1, dataLen));
        memcpy(R->o, o.data(), sizeof(float) * (size_t)dataLen);
    }
    return 0;
}

void sref_free_result(sref_result* R) {
    free(R->xx); free(R->m); free(R->logw); free(R->w_ceil); free(R->z_p); free(R->z); free(R->o);
    free(R->pcm);
    memset(R, 0, sizeof(*R));
}

void sref_destroy(void* h) {
    RefModel* M = (RefModel*)h;
    if (!M) return;
    delete M->enc; delete M->dp; delete M->flow; delete M->dec;
    delete M;
}

// ---------------------------------------------------------------------------------------------
// Op-level known-answer entry points.  `rec` is a float record in the .bin format of the
// respective constructor; x is time-major [T][Cin]; y is malloc'd time-major.
// ---------------------------------------------------------------------------------------------
static float* ret_tm(const MatrixXf& y, int32_t* To, int32_t* Co) {
    *To = (int32_t)y.rows(); *Co = (int32_t)y.cols();
    return dump_tm(y);
}

// nn_conv1d.cpp:118-199.  mode 0: blob ctor (pad/dil from record); mode 1: override ctor
// (pad, dil, sep given) as used by WN.cpp:42 / DDSConv.cpp:38-39.
float* sref_conv1d(const float* rec, const float* x, int32_t T, int32_t Cin, int32_t mode,
                   int32_t pad, int32_t dil, int32_t sep, int32_t* To, int32_t* Co) {
    int32_t off = 0;
    nn_conv1d* c = mode == 0 ? new nn_conv1d(const_cast<float*>(rec), off)
                             : new nn_conv1d(const_cast<float*>(rec), off, pad, dil, sep);
    MatrixXf y = c->forward(load_tm(x, T, Cin));
    delete c;
    return ret_tm(y, To, Co);
}

// nn_conv1d_transposed.cpp:106-150 with the (stride, padding) override ctor the generators use.
float* sref_conv1d_transposed(const float* rec, const float* x, int32_t T, int32_t Cin,
                              int32_t stride, int32_t pad, int32_t* To, int32_t* Co) {
    int32_t off = 0;
    nn_conv1d_transposed* c = new nn_conv1d_transposed(const_cast<float*>(rec), off, stride, pad);
    MatrixXf y = c->forward(load_tm(x, T, Cin));
    delete c;
    return ret_tm(y, To, Co);
}

// nn_layer_norm.cpp:65-86
float* sref_layer_norm(const float* rec, const float* x, int32_t T, int32_t C, int32_t* To, int32_t* Co) {
    int32_t off = 0;
    nn_layer_norm ln(const_cast<float*>(rec), off);
    MatrixXf y = ln.forward(load_tm(x, T, C));
    return ret_tm(y, To, Co);
}

// iStft.cpp:46-124 (16,4,16); mag/phase time-major [frames][9]; returns [1][(frames-1)*4]
float* sref_istft(const float* mag, const float* phase, int32_t frames, int32_t bins, int32_t* To, int32_t* Co) {
    iStft st(16, 4, 16);
    MatrixXf y = st.forward(load_tm(mag, frames, bins), load_tm(phase, frames, bins));
    return ret_tm(y, To, Co);
}

// pqmf.cpp:39-115; x time-major [T][4] -> [4T][1]
float* sref_pqmf(const float* x, int32_t T, int32_t* To, int32_t* Co) {
    pqmf p(4);
    MatrixXf y = p.forward(load_tm(x, T, 4));
    return ret_tm(y, To, Co);
}

// WN.cpp:100-149; rec = WN record; g nullable ([1][gin])
float* sref_wn(const float* rec, int32_t isMS, const float* x, int32_t T, int32_t C, const float* g,
               int32_t gin, int32_t* To, int32_t* Co) {
    int32_t off = 0;
    WN wn(const_cast<float*>(rec), off, 1, isMS);
    MatrixXf gm;
    if (isMS) gm = load_tm(g, 1, gin);
    MatrixXf y = wn.forward(load_tm(x, T, C), gm);
    return ret_tm(y, To, Co);
}

// ResBlock1.cpp:55-69
float* sref_resblock1(const float* rec, const float* x, int32_t T, int32_t C, int32_t* To, int32_t* Co) {
    int32_t off = 0;
    ResBlock1 rb(const_cast<float*>(rec), off);
    MatrixXf y = rb.forward(load_tm(x, T, C));
    return ret_tm(y, To, Co);
}

// multi_head_attention.cpp:106-121
float* sref_mha(const float* rec, const float* x, int32_t T, int32_t C, int32_t* To, int32_t* Co) {
    int32_t off = 0;
    multi_head_attention mha(const_cast<float*>(rec), off);
    MatrixXf xm = load_tm(x, T, C);
    MatrixXf y = mha.forward(xm, xm);
    return ret_tm(y, To, Co);
}

// elementwise: 0 tanh (nn_tanh.cpp), 1 gelu (nn_gelu.cpp)
float* sref_eltwise(int32_t which, const float* x, int32_t n) {
    MatrixXf xm = load_tm(x, n, 1);
    MatrixXf y = which == 0 ? nn_tanh(xm) : nn_gelu(xm);
    int32_t a, b;
    return ret_tm(y, &a, &b);
}

void sref_free(void* p) { free(p); }

}  // extern "C"
