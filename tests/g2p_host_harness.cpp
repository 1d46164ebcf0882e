// tests/g2p_host_harness.cpp — TEST INFRASTRUCTURE: runs the per-thread phases of the GPU g2p kernel
// (summertts_b200/csrc/g2p_phases.hpp, the very functions g2p_words_kernel calls between its barriers) thread by thread
// on the CPU, with the kernel's step loop mirrored statement by statement (summertts_b200/csrc/g2p.cuh), so that the
// kernel's indexing is checked against the oracle without a GPU (tests/test_g2p.py compiles this with g++).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../summertts_b200/csrc/g2p_phases.hpp"

using namespace stts;

namespace {
struct Mat { int r = 0, c = 0; const float* p = nullptr; };
Mat mat(const float* s, int64_t& cur) { Mat m; m.r = (int)s[cur++]; m.c = (int)s[cur++]; m.p = s + cur; cur += (int64_t)m.r * m.c; return m; }
Mat vec(const float* s, int64_t& cur) { Mat m; m.r = 1; m.c = (int)s[cur++]; m.p = s + cur; cur += m.c; return m; }
}  // namespace

extern "C" int g2p_host_predict(const float* sec, int32_t n_words, const char* letters, const int32_t* offsets, int32_t* preds,
                                int32_t* npreds, float* enc_hidden, float* first_logits) {
    int64_t cur = 0;
    const Mat eemb = mat(sec, cur), ewih = mat(sec, cur), ewhh = mat(sec, cur), ebih = vec(sec, cur), ebhh = vec(sec, cur);
    const Mat demb = mat(sec, cur), dwih = mat(sec, cur), dwhh = mat(sec, cur), dbih = vec(sec, cur), dbhh = vec(sec, cur);
    const Mat fcw = mat(sec, cur), fcb = vec(sec, cur);
    const int H = ewhh.c, H3 = 3 * H, V = demb.r;
    const int threads = ((std::max(H3, G2P_WPC * V) + 31) / 32) * 32;
    std::vector<float> etab((size_t)eemb.r * H3), dtab((size_t)demb.r * H3);
    for (int v = 0; v < eemb.r; ++v)                                   // g2p_table_kernel<<<V, threads>>>
        for (int i = 0; i < threads; ++i)
            if (i < H3) g2p_table_phase(v, i, eemb.r, eemb.c, H3, eemb.p, ewih.p, ebih.p, etab.data());
    for (int v = 0; v < demb.r; ++v)
        for (int i = 0; i < threads; ++i)
            if (i < H3) g2p_table_phase(v, i, demb.r, demb.c, H3, demb.p, dwih.p, dbih.p, dtab.data());
    std::vector<int32_t> order(n_words);
    for (int w = 0; w < n_words; ++w) order[w] = w;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b]; });
    memset(preds, 0, sizeof(int32_t) * (size_t)n_words * G2P_MAX_STEPS);

    const int ctas = (n_words + G2P_WPC - 1) / G2P_WPC;
    for (int cta = 0; cta < ctas; ++cta) {                             // g2p_words_kernel<<<ctas, threads>>>, one CTA at a time
        std::vector<float> hs((size_t)H * G2P_WPC, 0.f), gs((size_t)G2P_WPC * H3, -7.f), lg((size_t)G2P_WPC * V, -7.f);
        int tok[G2P_WPC], npred[G2P_WPC], wlen[G2P_WPC], woff[G2P_WPC], widx[G2P_WPC];
        for (int tid = 0; tid < G2P_WPC; ++tid) {
            const int s = cta * G2P_WPC + tid;
            const bool has = s < n_words;
            const int w = has ? order[s] : 0;
            widx[tid] = w;
            woff[tid] = has ? offsets[w] : 0;
            wlen[tid] = has ? offsets[w + 1] - offsets[w] : -1;
            npred[tid] = 0;
        }
        int maxlen = -1;
        for (int w = 0; w < G2P_WPC; ++w) maxlen = std::max(maxlen, wlen[w]);
        for (int t = 0; t <= maxlen; ++t) {
            for (int tid = 0; tid < G2P_WPC; ++tid) {
                const int L = wlen[tid];
                tok[tid] = (L < 0 || t > L) ? -1 : (t == L ? G2P_EOS_IN : g2p_letter_id((uint8_t)letters[woff[tid] + t]));
            }
            for (int tid = 0; tid < threads; ++tid)
                if (tid < H3) g2p_gates_phase(tid, H, ewhh.p, ebhh.p, etab.data(), tok, hs.data(), gs.data());
            for (int tid = 0; tid < threads; ++tid)
                if (tid < H) g2p_update_phase(tid, H, etab.data(), tok, gs.data(), hs.data());
        }
        if (enc_hidden)
            for (int i = 0; i < H * G2P_WPC; ++i) {
                const int j = i / G2P_WPC, w = i - j * G2P_WPC;
                if (wlen[w] >= 0) enc_hidden[(int64_t)widx[w] * H + j] = hs[i];
            }
        for (int tid = 0; tid < G2P_WPC; ++tid) tok[tid] = wlen[tid] < 0 ? -1 : G2P_BOS_OUT;
        for (int step = 0; step < G2P_MAX_STEPS; ++step) {
            bool any = false;
            for (int w = 0; w < G2P_WPC; ++w) any |= tok[w] >= 0;
            if (!any) break;
            for (int tid = 0; tid < threads; ++tid)
                if (tid < H3) g2p_gates_phase(tid, H, dwhh.p, dbhh.p, dtab.data(), tok, hs.data(), gs.data());
            for (int tid = 0; tid < threads; ++tid)
                if (tid < H) g2p_update_phase(tid, H, dtab.data(), tok, gs.data(), hs.data());
            for (int tid = 0; tid < threads; ++tid)
                if (tid < G2P_WPC * V) g2p_logits_phase(tid, H, V, fcw.p, fcb.p, tok, hs.data(), lg.data());
            if (step == 0 && first_logits)
                for (int i = 0; i < G2P_WPC * V; ++i) {
                    const int w = i / V;
                    if (wlen[w] >= 0) first_logits[(int64_t)widx[w] * V + (i - w * V)] = lg[i];
                }
            for (int tid = 0; tid < G2P_WPC; ++tid)
                if (tok[tid] >= 0) g2p_pick_phase(tid, V, lg.data(), tok, npred, preds + (int64_t)widx[tid] * G2P_MAX_STEPS);
        }
        for (int tid = 0; tid < G2P_WPC; ++tid)
            if (wlen[tid] >= 0) npreds[widx[tid]] = npred[tid];
    }
    return 0;
}
