// oracle/ref_g2p.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Driver around the reference's UNMODIFIED English g2p object (/root/reference/src/engipa/EnglishText2Id.cpp, compiled
// in place by oracle/Makefile) for the batched GPU g2p of summertts_b200/csrc/g2p.cuh:
//   * sref_g2p_ipa_ids : the reference's own EnglishText2Id::getIPAId(text)                      (:456-609), unmodified;
//   * sref_g2p_word    : the out-of-vocabulary branch of getIPAId (:496-545) restated over the reference's own
//                        gru() / gru_cell() (:270-313; free functions with external linkage in that object), returning
//                        what the reference keeps internal: the phone ids `preds`, the encoder state and the logits
//                        of the first decoder step.
// tests/test_g2p.py closes the loop: preds -> IPA ids through the frontend's tables must equal sref_g2p_ipa_ids(word)
// for words outside the dictionary, so the restated loop is pinned by the unmodified getIPAId.

#include <Eigen/Dense>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "EnglishText2Id.h"

using Eigen::Map;
using Eigen::MatrixXf;

// defined in the reference object EnglishText2Id.o (EnglishText2Id.cpp:270, :296)
MatrixXf gru_cell(const MatrixXf& x, const MatrixXf& h, const MatrixXf& w_ih, const MatrixXf& w_hh, const MatrixXf& b_ih, const MatrixXf& b_hh);
MatrixXf gru(const MatrixXf& x, int32_t steps, const MatrixXf& w_ih, const MatrixXf& w_hh, const MatrixXf& b_ih, const MatrixXf& b_hh,
             const MatrixXf& h0);

namespace {

struct RefG2p {
    std::vector<float> blob;          // the reference keeps Maps into the caller's buffer: own a copy
    EnglishText2Id* front = nullptr;  // unmodified reference object
    int32_t consumed = 0;
    MatrixXf enc_emb, enc_w_ih, enc_w_hh, enc_b_ih, enc_b_hh, dec_emb, dec_w_ih, dec_w_hh, dec_b_ih, dec_b_hh, fc_w, fc_b;
};

// record walk of the constructor, EnglishText2Id.cpp:73-126
MatrixXf take_mat(float* p, int32_t& cur) {
    const int32_t x = (int32_t)p[cur++], y = (int32_t)p[cur++];
    MatrixXf m = Map<MatrixXf>(p + cur, x, y);
    cur += x * y;
    return m;
}
MatrixXf take_vec(float* p, int32_t& cur) {
    const int32_t x = (int32_t)p[cur++];
    MatrixXf m = Map<MatrixXf>(p + cur, 1, x);
    cur += x;
    return m;
}

}  // namespace

extern "C" {

void* sref_g2p_create(const float* section, int64_t n) {
    RefG2p* g = new RefG2p();
    g->blob.assign(section, section + n);
    float* p = g->blob.data();
    int32_t off = 0;
    g->front = new EnglishText2Id(p, off);
    g->consumed = off;
    int32_t cur = 0;
    g->enc_emb = take_mat(p, cur); g->enc_w_ih = take_mat(p, cur); g->enc_w_hh = take_mat(p, cur);
    g->enc_b_ih = take_vec(p, cur); g->enc_b_hh = take_vec(p, cur);
    g->dec_emb = take_mat(p, cur); g->dec_w_ih = take_mat(p, cur); g->dec_w_hh = take_mat(p, cur);
    g->dec_b_ih = take_vec(p, cur); g->dec_b_hh = take_vec(p, cur);
    g->fc_w = take_mat(p, cur); g->fc_b = take_vec(p, cur);
    return g;
}

int32_t sref_g2p_consumed(void* h) { return ((RefG2p*)h)->consumed; }
int32_t sref_g2p_hidden(void* h) { return (int32_t)((RefG2p*)h)->enc_w_hh.cols(); }
int32_t sref_g2p_phones(void* h) { return (int32_t)((RefG2p*)h)->fc_w.rows(); }

void sref_g2p_destroy(void* h) {
    RefG2p* g = (RefG2p*)h;
    delete g->front;
    delete g;
}

// The reference's own text -> ids (dictionary, GRU for unknown words, IPA symbol table).  Returns the count; ids beyond cap are dropped.
int32_t sref_g2p_ipa_ids(void* h, const char* text, int32_t* out, int32_t cap) {
    std::vector<int> v = ((RefG2p*)h)->front->getIPAId(std::string(text));
    for (size_t i = 0; i < v.size() && (int32_t)i < cap; ++i) out[i] = v[i];
    return (int32_t)v.size();
}

// EnglishText2Id.cpp:496-545 for ONE lower-cased word: returns the number of predicted phones; preds[20], hidden[H], logits0[V].
int32_t sref_g2p_word(void* h, const char* word, int32_t* preds, float* hidden, float* logits0) {
    RefG2p* g = (RefG2p*)h;
    const int32_t wordSize = (int32_t)strlen(word);
    MatrixXf enc = MatrixXf::Zero(wordSize + 1, g->enc_emb.cols());
    for (int32_t l = 0; l < wordSize; ++l) {
        const char c = word[l];
        enc.row(l) = g->enc_emb.row((c >= 'a' && c <= 'z') ? 3 + (c - 'a') : 1);   // char2Id_ (:130-158), <unk> = 1 (:508)
    }
    enc.row(wordSize) = g->enc_emb.row(2);                                           // </s> (:513)
    MatrixXf h0 = MatrixXf::Zero(1, g->enc_w_hh.cols());
    MatrixXf out = gru(enc, wordSize + 1, g->enc_w_ih, g->enc_w_hh, g->enc_b_ih, g->enc_b_hh, h0);
    MatrixXf hh = out.row(out.rows() - 1);
    if (hidden) memcpy(hidden, hh.data(), sizeof(float) * hh.cols());
    MatrixXf dec = g->dec_emb.row(2);
    int32_t n = 0;
    for (int32_t i = 0; i < 20; ++i) {
        hh = gru_cell(dec, hh, g->dec_w_ih, g->dec_w_hh, g->dec_b_ih, g->dec_b_hh);
        MatrixXf logits = (hh * g->fc_w.transpose()) + g->fc_b;
        if (i == 0 && logits0) memcpy(logits0, logits.data(), sizeof(float) * logits.cols());
        MatrixXf::Index r, c;
        logits.maxCoeff(&r, &c);
        if (c == 3) break;
        preds[n++] = (int32_t)c;
        dec = g->dec_emb.row(c);
    }
    return n;
}

}  // extern "C"
