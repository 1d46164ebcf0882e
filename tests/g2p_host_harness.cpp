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

// ---- the cluster-resident kernel (g2p_cluster_kernel): 8 emulated CTAs per cluster, each with its own "shared memory"; a phase
// runs for every rank before the next one starts (the cluster barrier), remote writes go straight into the peers' arrays.
extern "C" int g2p_host_predict_cluster(const float* sec, int32_t n_words, const char* letters, const int32_t* offsets, int32_t* preds,
                                        int32_t* npreds, float* enc_hidden, float* first_logits, int32_t nclusters) {
    int64_t cur0 = 0;
    const Mat eemb = mat(sec, cur0), ewih = mat(sec, cur0), ewhh = mat(sec, cur0), ebih = vec(sec, cur0), ebhh = vec(sec, cur0);
    const Mat demb = mat(sec, cur0), dwih = mat(sec, cur0), dwhh = mat(sec, cur0), dbih = vec(sec, cur0), dbhh = vec(sec, cur0);
    const Mat fcw = mat(sec, cur0), fcb = vec(sec, cur0);
    const int H = ewhh.c, H3 = 3 * H, V = demb.r;
    if (H % G2P_CL != 0 || H % 8 != 0) return -1;
    const G2pClDims d = g2p_cl_dims(H, V);
    const int HW = H * G2P_WG, nt = 256;
    std::vector<float> etab((size_t)eemb.r * H3), dtab((size_t)demb.r * H3);
    for (int v = 0; v < eemb.r; ++v)
        for (int i = 0; i < H3; ++i) g2p_table_phase(v, i, eemb.r, eemb.c, H3, eemb.p, ewih.p, ebih.p, etab.data());
    for (int v = 0; v < demb.r; ++v)
        for (int i = 0; i < H3; ++i) g2p_table_phase(v, i, demb.r, demb.c, H3, demb.p, dwih.p, dbih.p, dtab.data());
    std::vector<int32_t> order(n_words);
    for (int w = 0; w < n_words; ++w) order[w] = w;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b]; });
    memset(preds, 0, sizeof(int32_t) * (size_t)n_words * G2P_MAX_STEPS);
    const int ngroups = (n_words + G2P_WG - 1) / G2P_WG;
    const int ncl = std::min(ngroups, nclusters);

    struct Cta {
        std::vector<float> sm;
        float *We, *Wd, *Fw, *hb, *gs, *lg;
        int tok[G2P_WG], npred[G2P_WG], wlen[G2P_WG], woff[G2P_WG], widx[G2P_WG];
    };
    for (int cid = 0; cid < ncl; ++cid) {
        Cta c[G2P_CL];
        for (int rank = 0; rank < G2P_CL; ++rank) {                    // carve-up and prologue of the kernel
            Cta& k = c[rank];
            k.sm.assign((size_t)g2p_cl_smem_floats(d), -7.f);
            k.We = k.sm.data(); k.Wd = k.We + H * d.R; k.Fw = k.Wd + H * d.R; k.hb = k.Fw + H * d.VS; k.gs = k.hb + 2 * HW; k.lg = k.gs + d.R * G2P_WG;
            if (k.lg + d.VS * G2P_CL * G2P_WG != k.sm.data() + k.sm.size()) return -2;
            for (int tid = 0; tid < nt; ++tid) {
                for (int i = tid; i < H * d.R; i += nt) { g2p_cl_load_w(i, d, rank, ewhh.p, k.We); g2p_cl_load_w(i, d, rank, dwhh.p, k.Wd); }
                for (int i = tid; i < H * d.VS; i += nt) g2p_cl_load_fc(i, d, rank, fcw.p, k.Fw);
            }
        }
        for (int g = cid; g < ngroups; g += ncl) {
            int maxlen = -1;
            for (int rank = 0; rank < G2P_CL; ++rank) {
                Cta& k = c[rank];
                for (int i = 0; i < HW; ++i) k.hb[i] = 0.f;
                for (int tid = 0; tid < G2P_WG; ++tid) {
                    const int s = g * G2P_WG + tid;
                    const bool has = s < n_words;
                    const int w = has ? order[s] : 0;
                    k.widx[tid] = w; k.woff[tid] = has ? offsets[w] : 0; k.wlen[tid] = has ? offsets[w + 1] - offsets[w] : -1; k.npred[tid] = 0;
                }
            }
            for (int w = 0; w < G2P_WG; ++w) maxlen = std::max(maxlen, c[0].wlen[w]);
            int cur = 0;
            auto step_cell = [&](bool enc) {
                for (int rank = 0; rank < G2P_CL; ++rank) {
                    Cta& k = c[rank];
                    for (int tid = 0; tid < nt; ++tid)
                        for (int o = tid; o < 2 * d.R; o += nt)
                            g2p_cl_gates_phase(o, d, rank, enc ? k.We : k.Wd, enc ? ebhh.p : dbhh.p, enc ? etab.data() : dtab.data(), k.tok, k.hb + cur * HW, k.gs);
                }
                for (int rank = 0; rank < G2P_CL; ++rank) {
                    Cta& k = c[rank];
                    float* hn[G2P_CL];
                    for (int rk = 0; rk < G2P_CL; ++rk) hn[rk] = c[rk].hb + (cur ^ 1) * HW;
                    for (int tid = 0; tid < nt; ++tid)
                        for (int i = tid; i < d.HS * G2P_WG; i += nt) g2p_cl_update_phase(i, d, rank, enc ? etab.data() : dtab.data(), k.tok, k.gs, k.hb + cur * HW, hn);
                }
                cur ^= 1;
            };
            for (int t = 0; t <= maxlen; ++t) {
                for (int rank = 0; rank < G2P_CL; ++rank)
                    for (int tid = 0; tid < G2P_WG; ++tid) {
                        Cta& k = c[rank];
                        const int L = k.wlen[tid];
                        k.tok[tid] = (L < 0 || t > L) ? -1 : (t == L ? G2P_EOS_IN : g2p_letter_id((uint8_t)letters[k.woff[tid] + t]));
                    }
                step_cell(true);
            }
            if (enc_hidden)
                for (int i = 0; i < HW; ++i) {
                    const int j = i / G2P_WG, w = i - j * G2P_WG;
                    if (c[0].wlen[w] >= 0) enc_hidden[(int64_t)c[0].widx[w] * H + j] = c[0].hb[cur * HW + i];
                }
            for (int rank = 0; rank < G2P_CL; ++rank)
                for (int tid = 0; tid < G2P_WG; ++tid) c[rank].tok[tid] = c[rank].wlen[tid] < 0 ? -1 : G2P_BOS_OUT;
            for (int step = 0; step < G2P_MAX_STEPS; ++step) {
                bool any = false;
                for (int w = 0; w < G2P_WG; ++w) any |= c[3].tok[w] >= 0;
                if (!any) break;
                step_cell(false);
                for (int rank = 0; rank < G2P_CL; ++rank) {
                    Cta& k = c[rank];
                    float* lgs[G2P_CL];
                    for (int rk = 0; rk < G2P_CL; ++rk) lgs[rk] = c[rk].lg;
                    for (int tid = 0; tid < nt; ++tid)
                        for (int i = tid; i < d.VS * G2P_WG; i += nt) g2p_cl_logits_phase(i, d, rank, k.Fw, fcb.p, k.tok, k.hb + cur * HW, lgs);
                }
                if (step == 0 && first_logits)
                    for (int i = 0; i < d.V * G2P_WG; ++i) {
                        const int cc = i / G2P_WG, w = i - cc * G2P_WG;
                        if (c[0].wlen[w] >= 0) first_logits[(int64_t)c[0].widx[w] * d.V + cc] = c[0].lg[i];
                    }
                for (int rank = 0; rank < G2P_CL; ++rank)
                    for (int tid = 0; tid < G2P_WG; ++tid) {
                        Cta& k = c[rank];
                        if (k.tok[tid] >= 0) g2p_cl_pick_phase(tid, d.V, k.lg, k.tok, k.npred, rank == 0 ? preds + (int64_t)k.widx[tid] * G2P_MAX_STEPS : nullptr);
                    }
                for (int rank = 1; rank < G2P_CL; ++rank)            // the kernel relies on every CTA holding the same tokens
                    for (int w = 0; w < G2P_WG; ++w)
                        if (c[rank].tok[w] != c[0].tok[w]) return -3;
            }
            for (int tid = 0; tid < G2P_WG; ++tid)
                if (c[0].wlen[tid] >= 0) npreds[c[0].widx[tid]] = c[0].npred[tid];
        }
    }
    return 0;
}
