// g2p_phases.hpp — the per-thread phases of the batched GRU grapheme-to-phoneme kernel (g2p.cuh).
//
// Replaces, for out-of-vocabulary English words, the Eigen GRU of the reference's host frontend:
//   gru_cell   /root/reference/src/engipa/EnglishText2Id.cpp:270-294
//   gru        :296-313          (encoder over the letters + </s>)
//   greedy decoder loop          :521-540  (<= 20 steps, argmax, stop at id 3)
//   letter -> id map             :496-513  ('a'..'z' -> 3..28, anything else -> <unk> = 1, </s> = 2)
//
// Every function below is the work of ONE thread `tid` between two block barriers; all indexing of the kernel lives
// here.  The file is plain C++ when STTS_HD is empty, so tests/g2p_host_harness.cpp runs the very same phases thread
// by thread on the CPU against the reference results (the GPU box is not needed to check the indexing).
//
// Memory layout (as stored in the `.bin`: Eigen column-major `Map<MatrixXf>(p, rows, cols)` -> element (r, c) at
// p[r + c * rows], EnglishText2Id.cpp:75-123):
//   whh  (3H, H)  -> whh[j * 3H + i]   = W_hh(i, j): consecutive threads i read consecutive floats (coalesced)
//   fcw  (V, H)   -> fcw[j * V + c]    = fc_w(c, j)
//   tab  [V_in][3H] = emb(v, :) . W_ih^T + b_ih, built once at create time by g2p_table_phase (the input half of a
//                     GRU cell depends only on the token, so it is a table lookup per step)
// Shared-memory state of a CTA that carries WPC words:
//   hs[j * WPC + w]     hidden state h_w[j]   (one 16-byte load gives h_j of all four words)
//   gs[w * 3H + i]      gate scratch: sigmoid(r | z pre-activation) for i < 2H, the hidden half of the n gate for i >= 2H
//   lg[w * V + c]       logits
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef STTS_HD
#define STTS_HD
#endif

namespace stts {

constexpr int G2P_WPC = 4;         // words per CTA: one pass over W_hh serves four hidden states
constexpr int G2P_MAX_STEPS = 20;  // EnglishText2Id.cpp:527
constexpr int G2P_EOS_IN = 2;      // "</s>" of the letter table (:513)
constexpr int G2P_BOS_OUT = 2;     // "<s>" of the phone table (:520)
constexpr int G2P_EOS_OUT = 3;     // "</s>" of the phone table: stops the decoder (:536)

STTS_HD inline float g2p_tanh(float x) {   // nn_tanh, src/nn_op/nn_tanh.cpp:6-21
    float a = expf(x), b = expf(-x);
    if (isinf(a)) a = 1e10f;
    if (isinf(b)) b = 1e10f;
    float d = a + b;
    if (d < 1e-8f) d = 1e-8f;
    return (a - b) / d;
}
STTS_HD inline float g2p_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }   // nn_sigmoid.cpp:3-7

// h_j of the CTA's four words: one 16-byte shared-memory load on the device (hs is 16-byte aligned)
STTS_HD inline void g2p_load_h(const float* hs, int j, float (&h)[G2P_WPC]) {
#ifdef __CUDA_ARCH__
    const float4 v = *reinterpret_cast<const float4*>(hs + j * G2P_WPC);
    h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
#else
    for (int w = 0; w < G2P_WPC; ++w) h[w] = hs[j * G2P_WPC + w];
#endif
}
static_assert(G2P_WPC == 4, "g2p_load_h reads one float4 per hidden unit");

STTS_HD inline int g2p_letter_id(uint8_t c) { return (c >= 'a' && c <= 'z') ? 3 + (int)(c - 'a') : 1; }

// tab[v][i] = b_ih[i] + sum_e emb(v, e) * W_ih(i, e)        (thread i of block v; the `x * w_ih^T + b_ih` of :274)
STTS_HD inline void g2p_table_phase(int v, int i, int V, int E, int H3, const float* emb, const float* wih, const float* bih, float* tab) {
    float acc = bih[i];
    for (int e = 0; e < E; ++e) acc = fmaf(emb[v + (int64_t)e * V], wih[i + (int64_t)e * H3], acc);
    tab[(int64_t)v * H3 + i] = acc;
}

// Phase A (thread i < 3H): hidden half of the cell for the CTA's words, `h * w_hh^T + b_hh` (:275), then the r | z gates (:283).
// tok[w] = row of `tab` for word w at this step, or -1 when the word takes no step.
STTS_HD inline void g2p_gates_phase(int i, int H, const float* whh, const float* bhh, const float* tab, const int* tok, const float* hs, float* gs) {
    const int H3 = 3 * H;
    float acc[G2P_WPC];
    const float b = bhh[i];
    for (int w = 0; w < G2P_WPC; ++w) acc[w] = b;
#pragma unroll 8
    for (int j = 0; j < H; ++j) {
        const float wv = whh[(int64_t)j * H3 + i];
        float h[G2P_WPC];
        g2p_load_h(hs, j, h);
        for (int w = 0; w < G2P_WPC; ++w) acc[w] = fmaf(h[w], wv, acc[w]);
    }
    for (int w = 0; w < G2P_WPC; ++w) {
        if (tok[w] < 0) continue;
        gs[w * H3 + i] = (i < 2 * H) ? g2p_sigmoid(tab[(int64_t)tok[w] * H3 + i] + acc[w]) : acc[w];
    }
}

// Phase B (thread i < H): n = tanh(n_ih + r * n_hh) (:288), h' = (1 - z) * n + z * h (:290).
STTS_HD inline void g2p_update_phase(int i, int H, const float* tab, const int* tok, const float* gs, float* hs) {
    const int H3 = 3 * H;
    for (int w = 0; w < G2P_WPC; ++w) {
        if (tok[w] < 0) continue;
        const float r = gs[w * H3 + i], z = gs[w * H3 + H + i];
        const float n = g2p_tanh(tab[(int64_t)tok[w] * H3 + 2 * H + i] + r * gs[w * H3 + 2 * H + i]);
        const float h = hs[i * G2P_WPC + w];
        hs[i * G2P_WPC + w] = (z * -1.0f + 1.0f) * n + z * h;
    }
}

// Phase C (thread t < WPC * V): logits = h * fc_w^T + fc_b (:534).
STTS_HD inline void g2p_logits_phase(int t, int H, int V, const float* fcw, const float* fcb, const int* tok, const float* hs, float* lg) {
    const int w = t / V, c = t - w * V;
    if (tok[w] < 0) return;
    float acc = fcb[c];
    for (int j = 0; j < H; ++j) acc = fmaf(hs[j * G2P_WPC + w], fcw[(int64_t)j * V + c], acc);
    lg[w * V + c] = acc;
}

// Phase D (thread w < WPC): first maximum of the logits (Eigen maxCoeff keeps the first, :537), stop at </s> (:539),
// else record the phone and feed it back (:542-543).  tok[w] = -1 retires the word.
STTS_HD inline void g2p_pick_phase(int w, int V, const float* lg, int* tok, int* npred, int32_t* preds /* [G2P_MAX_STEPS] of word w */) {
    if (tok[w] < 0) return;
    int best = 0;
    float bv = lg[w * V];
    for (int c = 1; c < V; ++c)
        if (lg[w * V + c] > bv) { bv = lg[w * V + c]; best = c; }
    if (best == G2P_EOS_OUT) { tok[w] = -1; return; }
    preds[npred[w]] = best;
    npred[w] += 1;
    tok[w] = (npred[w] >= G2P_MAX_STEPS) ? -1 : best;
}

// ---------------------------------------------------------------------------------------------------------------------
// Cluster-resident variant (g2p_cluster_kernel): a thread-block cluster of G2P_CL = 8 CTAs keeps BOTH recurrent matrices
// in shared memory for the whole launch — CTA `rank` owns hidden units [rank*HS, (rank+1)*HS), i.e. the R = 3*HS rows
// {g*H + rank*HS + u} of W_hh (96 rows x 256 x 4 B = 96 KB per matrix at H = 256) plus VS = ceil(V / 8) rows of fc_w —
// and the CTAs exchange the new hidden-state slices (and the logits slices) by writing them into every peer's shared
// memory (DSMEM) before a cluster barrier.  No weight byte is re-read from L2 after the prologue; a step costs one
// shared-memory GEMV slice + one cluster.sync.  A cluster carries G2P_WG = 8 words.
//   Ws[j * R + r]        slice of W_hh, r fastest (conflict-free across threads), row(r) = (r / HS) * H + rank*HS + r % HS
//   Fw[j * VS + cc]      slice of fc_w, class c = rank * VS + cc (zero beyond V)
//   hb[2][j * WG + w]    full hidden state of the cluster's words, double-buffered: peers write buffer cur^1 while
//                        buffer cur is still being read
//   gs[r * WG + w]       gate scratch of the CTA's own rows
//   lg[c * WG + w]       logits of ALL classes (every CTA receives every slice and takes the argmax redundantly, so the
//                        next token is known everywhere without a second exchange)
// The accumulation order per output is the same as in the streaming kernel (bias, then j = 0..H-1 by fmaf): both
// kernels produce bit-identical hidden states and logits.
constexpr int G2P_CL = 8;
constexpr int G2P_WG = 8;
struct G2pClDims {
    int H, HS, R, V, VS;
};
STTS_HD inline G2pClDims g2p_cl_dims(int H, int V) { return G2pClDims{H, H / G2P_CL, 3 * (H / G2P_CL), V, (V + G2P_CL - 1) / G2P_CL}; }
// floats of shared memory per CTA (the 5 * WG ints of word state come on top)
STTS_HD inline int64_t g2p_cl_smem_floats(const G2pClDims& d) {
    return 2 * (int64_t)d.H * d.R + (int64_t)d.H * d.VS + 2 * (int64_t)d.H * G2P_WG + (int64_t)d.R * G2P_WG + (int64_t)d.VS * G2P_CL * G2P_WG;
}
STTS_HD inline int g2p_cl_row(const G2pClDims& d, int rank, int r) { return (r / d.HS) * d.H + rank * d.HS + (r % d.HS); }

// prologue (item idx < H * R / idx < H * VS): the CTA's slices of W_hh and fc_w
STTS_HD inline void g2p_cl_load_w(int idx, const G2pClDims& d, int rank, const float* whh, float* Ws) {
    const int j = idx / d.R, r = idx - j * d.R;
    Ws[idx] = whh[(int64_t)j * 3 * d.H + g2p_cl_row(d, rank, r)];
}
STTS_HD inline void g2p_cl_load_fc(int idx, const G2pClDims& d, int rank, const float* fcw, float* Fw) {
    const int j = idx / d.VS, cc = idx - j * d.VS, c = rank * d.VS + cc;
    Fw[idx] = c < d.V ? fcw[(int64_t)j * d.V + c] : 0.f;
}

// Phase A (item o < 2 * R: row r = o % R of the slice, word quad wq = o / R)
STTS_HD inline void g2p_cl_gates_phase(int o, const G2pClDims& d, int rank, const float* Ws, const float* bhh, const float* tab, const int* tok,
                                       const float* hc, float* gs) {
    const int r = o % d.R, wq = o / d.R, row = g2p_cl_row(d, rank, r), H3 = 3 * d.H;
    float ih[4], acc[4];
    const float b = bhh[row];
    for (int q = 0; q < 4; ++q) {
        const int t = tok[wq * 4 + q];
        ih[q] = t >= 0 ? tab[(int64_t)t * H3 + row] : 0.f;      // issued before the GEMV loop: the L2 latency hides behind it
        acc[q] = b;
    }
#pragma unroll 8
    for (int j = 0; j < d.H; ++j) {
        const float wv = Ws[j * d.R + r];
        float h[4];
        g2p_load_h(hc + wq * 4, j * (G2P_WG / 4), h);             // hc[j * WG + wq * 4 .. + 3]
        for (int q = 0; q < 4; ++q) acc[q] = fmaf(h[q], wv, acc[q]);
    }
    for (int q = 0; q < 4; ++q) {
        const int w = wq * 4 + q;
        if (tok[w] < 0) continue;
        gs[r * G2P_WG + w] = (r < 2 * d.HS) ? g2p_sigmoid(ih[q] + acc[q]) : acc[q];
    }
}

// Phase B (item idx < HS * WG: own hidden unit u = idx / WG, word w = idx % WG): the new h of unit rank*HS + u goes into buffer
// cur^1 of EVERY CTA of the cluster (hn[rk] = that buffer in CTA rk); a word that takes no step carries its state over.
STTS_HD inline void g2p_cl_update_phase(int idx, const G2pClDims& d, int rank, const float* tab, const int* tok, const float* gs, const float* hc,
                                        float* const* hn) {
    const int u = idx / G2P_WG, w = idx - u * G2P_WG, U = rank * d.HS + u, H3 = 3 * d.H;
    float hv = hc[U * G2P_WG + w];
    if (tok[w] >= 0) {
        const float r = gs[u * G2P_WG + w], z = gs[(d.HS + u) * G2P_WG + w];
        const float n = g2p_tanh(tab[(int64_t)tok[w] * H3 + 2 * d.H + U] + r * gs[(2 * d.HS + u) * G2P_WG + w]);
        hv = (z * -1.0f + 1.0f) * n + z * hv;
    }
    for (int rk = 0; rk < G2P_CL; ++rk) hn[rk][U * G2P_WG + w] = hv;
}

// Phase C (item idx < VS * WG: class slot cc = idx / WG, word w): the CTA's slice of the logits, written to every CTA
STTS_HD inline void g2p_cl_logits_phase(int idx, const G2pClDims& d, int rank, const float* Fw, const float* fcb, const int* tok, const float* hc,
                                        float* const* lgs) {
    const int cc = idx / G2P_WG, w = idx - cc * G2P_WG, c = rank * d.VS + cc;
    if (c >= d.V || tok[w] < 0) return;
    float acc = fcb[c];
    for (int j = 0; j < d.H; ++j) acc = fmaf(hc[j * G2P_WG + w], Fw[j * d.VS + cc], acc);
    for (int rk = 0; rk < G2P_CL; ++rk) lgs[rk][c * G2P_WG + w] = acc;
}

// Phase D (thread w < WG, in every CTA): as g2p_pick_phase over lg[c * WG + w]; only rank 0 passes `preds` (global memory)
STTS_HD inline void g2p_cl_pick_phase(int w, int V, const float* lg, int* tok, int* npred, int32_t* preds) {
    if (tok[w] < 0) return;
    int best = 0;
    float bv = lg[w];
    for (int c = 1; c < V; ++c)
        if (lg[c * G2P_WG + w] > bv) { bv = lg[c * G2P_WG + w]; best = c; }
    if (best == G2P_EOS_OUT) { tok[w] = -1; return; }
    if (preds) preds[npred[w]] = best;
    npred[w] += 1;
    tok[w] = (npred[w] >= G2P_MAX_STEPS) ? -1 : best;
}

}  // namespace stts
