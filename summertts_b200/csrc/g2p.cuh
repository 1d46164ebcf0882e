// g2p.cuh — batched GRU grapheme-to-phoneme on the GPU: the out-of-vocabulary branch of the
// reference's English frontend, /root/reference/src/engipa/EnglishText2Id.cpp:496-540 (encoder GRU over the letters,
// greedy GRU decoder, <= 20 phones), for MANY words at once.  The reference runs it one word at a time on Eigen
// (two 768x256 GEMVs per step, ~30 steps per word); at 60 000x real time for the acoustic model that serial loop is
// what a batch of English utterances with unseen words waits for.
//
// Included at the end of engine.cu (uses its CUDA_CHECK / guard / error types).  All per-thread work is in
// g2p_phases.hpp, shared with the CPU harness of the tests; this file holds the step loop, the launch and the C ABI.
//
// Kernel shape: one CTA per G2P_WPC = 4 words (sorted by length so a CTA's words finish together), 3H threads
// (768 for the shipped model).  Per step every thread streams its column of W_hh (H floats, coalesced across the CTA,
// L2-resident: 786 KB per matrix) against the four hidden states held in shared memory — one weight load feeds four
// FMAs.  The input half of each cell is a table lookup (emb . W_ih^T + b_ih per token, built once at create time), so a
// step is one GEMV, not two.  Bound: L2 -> SM bandwidth of the W_hh stream (786 KB per step per CTA).
#pragma once
#define STTS_HD __host__ __device__
#include "g2p_phases.hpp"

namespace stts {

struct G2pDev {
    int H = 0, E = 0, Vin = 0, Vout = 0;
    const float *enc_tab = nullptr, *enc_whh = nullptr, *enc_bhh = nullptr;   // [Vin][3H], (3H, H) col-major, [3H]
    const float *dec_tab = nullptr, *dec_whh = nullptr, *dec_bhh = nullptr;   // [Vout][3H], ...
    const float *fcw = nullptr, *fcb = nullptr;                               // (Vout, H) col-major, [Vout]
};
struct G2pBatch {
    int n_words = 0;
    const uint8_t* letters = nullptr;   // lower-cased words, concatenated
    const int32_t* offsets = nullptr;   // [n_words + 1]
    const int32_t* order = nullptr;     // word indices sorted by length (descending)
    int32_t* preds = nullptr;           // [n_words][G2P_MAX_STEPS]
    int32_t* npreds = nullptr;          // [n_words]
    float* enc_hidden = nullptr;        // optional [n_words][H]: encoder state after </s>   (EnglishText2Id.cpp:519)
    float* first_logits = nullptr;      // optional [n_words][Vout]: logits of decoder step 0 (:534)
};

__global__ void g2p_table_kernel(int V, int E, int H3, const float* emb, const float* wih, const float* bih, float* tab) {
    const int i = threadIdx.x;
    if (i < H3) g2p_table_phase(blockIdx.x, i, V, E, H3, emb, wih, bih, tab);
}

__global__ void __launch_bounds__(1024) g2p_words_kernel(G2pDev m, G2pBatch b) {
    extern __shared__ __align__(16) float g2p_sm[];
    const int H = m.H, H3 = 3 * m.H, V = m.Vout;
    float* hs = g2p_sm;                          // [H][WPC]
    float* gs = hs + H * G2P_WPC;                // [WPC][3H]
    float* lg = gs + G2P_WPC * H3;               // [WPC][V]
    int* tok = (int*)(lg + G2P_WPC * V);         // [WPC] token of the current step, -1 = idle
    int* npred = tok + G2P_WPC;                  // [WPC]
    int* wlen = npred + G2P_WPC;                 // [WPC] letters of the word, -1 = no word in this slot
    int* woff = wlen + G2P_WPC;                  // [WPC]
    int* widx = woff + G2P_WPC;                  // [WPC] index of the word in the caller's order
    const int tid = threadIdx.x;

    for (int i = tid; i < H * G2P_WPC; i += blockDim.x) hs[i] = 0.f;    // h0 = 0 (:515)
    if (tid < G2P_WPC) {
        const int s = blockIdx.x * G2P_WPC + tid;
        const bool has = s < b.n_words;
        const int w = has ? b.order[s] : 0;
        widx[tid] = w;
        woff[tid] = has ? b.offsets[w] : 0;
        wlen[tid] = has ? b.offsets[w + 1] - b.offsets[w] : -1;
        npred[tid] = 0;
    }
    __syncthreads();
    int maxlen = -1;
#pragma unroll
    for (int w = 0; w < G2P_WPC; ++w) maxlen = max(maxlen, wlen[w]);

    // encoder: the word's letters, then </s> (:498-516)
    for (int t = 0; t <= maxlen; ++t) {
        if (tid < G2P_WPC) {
            const int L = wlen[tid];
            tok[tid] = (L < 0 || t > L) ? -1 : (t == L ? G2P_EOS_IN : g2p_letter_id(b.letters[woff[tid] + t]));
        }
        __syncthreads();
        if (tid < H3) g2p_gates_phase(tid, H, m.enc_whh, m.enc_bhh, m.enc_tab, tok, hs, gs);
        __syncthreads();
        if (tid < H) g2p_update_phase(tid, H, m.enc_tab, tok, gs, hs);
        __syncthreads();
    }
    if (b.enc_hidden)
        for (int i = tid; i < H * G2P_WPC; i += blockDim.x) {
            const int j = i / G2P_WPC, w = i - j * G2P_WPC;
            if (wlen[w] >= 0) b.enc_hidden[(int64_t)widx[w] * H + j] = hs[i];
        }

    // greedy decoder (:520-545)
    if (tid < G2P_WPC) tok[tid] = wlen[tid] < 0 ? -1 : G2P_BOS_OUT;
    __syncthreads();
    for (int step = 0; step < G2P_MAX_STEPS; ++step) {
        bool any = false;
#pragma unroll
        for (int w = 0; w < G2P_WPC; ++w) any |= tok[w] >= 0;      // same shared-memory words for every thread: uniform
        if (!any) break;
        if (tid < H3) g2p_gates_phase(tid, H, m.dec_whh, m.dec_bhh, m.dec_tab, tok, hs, gs);
        __syncthreads();
        if (tid < H) g2p_update_phase(tid, H, m.dec_tab, tok, gs, hs);
        __syncthreads();
        if (tid < G2P_WPC * V) g2p_logits_phase(tid, H, V, m.fcw, m.fcb, tok, hs, lg);
        __syncthreads();
        if (step == 0 && b.first_logits)
            for (int i = tid; i < G2P_WPC * V; i += blockDim.x) {
                const int w = i / V;
                if (wlen[w] >= 0) b.first_logits[(int64_t)widx[w] * V + (i - w * V)] = lg[i];
            }
        if (tid < G2P_WPC && tok[tid] >= 0) g2p_pick_phase(tid, V, lg, tok, npred, b.preds + (int64_t)widx[tid] * G2P_MAX_STEPS);
        __syncthreads();
    }
    if (tid < G2P_WPC && wlen[tid] >= 0) b.npreds[widx[tid]] = npred[tid];
}

}  // namespace stts

struct stts_g2p {
    int device = 0;
    cudaStream_t stream = nullptr;
    stts::G2pDev m;
    std::vector<void*> owned;
    uint8_t* d_letters = nullptr;
    int32_t *d_offsets = nullptr, *d_order = nullptr, *d_preds = nullptr, *d_npreds = nullptr;
    float *d_hidden = nullptr, *d_logits = nullptr;
    int64_t capLetters = 0, capWords = 0;
    int64_t launches = 0;
    size_t smem = 0;
    int threads = 0;

    template <typename T>
    T* dalloc(size_t n) {
        T* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        owned.push_back(p);
        return p;
    }
    const float* upload(const float* h, size_t n) {
        float* d = dalloc<float>(n);
        CUDA_CHECK(cudaMemcpyAsync(d, h, n * 4, cudaMemcpyHostToDevice, stream));
        return d;
    }
    void release(void* p) {
        if (!p) return;
        owned.erase(std::remove(owned.begin(), owned.end(), p), owned.end());
        cudaFree(p);
    }
    ~stts_g2p() {
        cudaSetDevice(device);
        for (void* p : owned) cudaFree(p);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace stts {

// gru section := enc_emb(V_in, E) enc_w_ih(3H, E) enc_w_hh(3H, H) enc_b_ih enc_b_hh dec_emb(V_out, E2) dec_w_ih(3H, E2)
//                dec_w_hh(3H, H) dec_b_ih dec_b_hh fc_w(V_out, H) fc_b        (EnglishText2Id.cpp:73-126; every matrix
//                preceded by rows, cols and every vector by its length, stored as floats)
struct G2pMat {
    int32_t r = 0, c = 0;
    const float* p = nullptr;
};
static G2pMat g2p_mat(Cursor& cur) {
    G2pMat a;
    a.r = cur.i(); a.c = cur.i();
    if (a.r <= 0 || a.c <= 0 || a.r > 65536 || a.c > 65536) throw FormatError("g2p section: implausible matrix shape");
    a.p = cur.f((int64_t)a.r * a.c);
    return a;
}
static G2pMat g2p_vec(Cursor& cur) {
    G2pMat a;
    a.r = 1; a.c = cur.i();
    if (a.c <= 0 || a.c > 65536) throw FormatError("g2p section: implausible vector length");
    a.p = cur.f(a.c);
    return a;
}

static void g2p_build(stts_g2p* g, const float* sec, int64_t n, int64_t* consumed) {
    Cursor cur{sec, n};
    const G2pMat eemb = g2p_mat(cur), ewih = g2p_mat(cur), ewhh = g2p_mat(cur), ebih = g2p_vec(cur), ebhh = g2p_vec(cur);
    const G2pMat demb = g2p_mat(cur), dwih = g2p_mat(cur), dwhh = g2p_mat(cur), dbih = g2p_vec(cur), dbhh = g2p_vec(cur);
    const G2pMat fcw = g2p_mat(cur), fcb = g2p_vec(cur);
    if (consumed) *consumed = cur.o;
    const int H = ewhh.c, H3 = 3 * H;
    if (ewhh.r != H3 || ewih.r != H3 || ebih.c != H3 || ebhh.c != H3 || dwhh.r != H3 || dwhh.c != H || dwih.r != H3 || dbih.c != H3 ||
        dbhh.c != H3 || ewih.c != eemb.c || dwih.c != demb.c || fcw.c != H || fcw.r != demb.r || fcb.c != fcw.r)
        throw FormatError("g2p section: inconsistent GRU shapes");
    if (eemb.r < 29) throw FormatError("g2p section: letter table smaller than the 29 ids the frontend emits");
    if (demb.r <= G2P_EOS_OUT) throw FormatError("g2p section: phone table without <s> / </s>");
    const int threads = ((std::max(H3, G2P_WPC * fcw.r) + 31) / 32) * 32;
    if (threads > 1024) throw Unsupported("g2p: 3 * hidden (or 4 * phones) exceeds one thread block");
    g->threads = threads;
    g->smem = ((size_t)H * G2P_WPC + (size_t)G2P_WPC * H3 + (size_t)G2P_WPC * fcw.r) * 4 + 5 * G2P_WPC * 4;
    G2pDev& m = g->m;
    m.H = H; m.E = eemb.c; m.Vin = eemb.r; m.Vout = demb.r;
    m.enc_whh = g->upload(ewhh.p, (size_t)H3 * H);
    m.enc_bhh = g->upload(ebhh.p, H3);
    m.dec_whh = g->upload(dwhh.p, (size_t)H3 * H);
    m.dec_bhh = g->upload(dbhh.p, H3);
    m.fcw = g->upload(fcw.p, (size_t)fcw.r * H);
    m.fcb = g->upload(fcb.p, fcb.c);
    // input halves of the two cells as per-token tables, computed on the device
    float* etab = g->dalloc<float>((size_t)eemb.r * H3);
    float* dtab = g->dalloc<float>((size_t)demb.r * H3);
    const float* t_eemb = g->upload(eemb.p, (size_t)eemb.r * eemb.c);
    const float* t_ewih = g->upload(ewih.p, (size_t)H3 * ewih.c);
    const float* t_ebih = g->upload(ebih.p, H3);
    const float* t_demb = g->upload(demb.p, (size_t)demb.r * demb.c);
    const float* t_dwih = g->upload(dwih.p, (size_t)H3 * dwih.c);
    const float* t_dbih = g->upload(dbih.p, H3);
    g2p_table_kernel<<<eemb.r, threads, 0, g->stream>>>(eemb.r, eemb.c, H3, t_eemb, t_ewih, t_ebih, etab);
    CUDA_CHECK(cudaGetLastError());
    g2p_table_kernel<<<demb.r, threads, 0, g->stream>>>(demb.r, demb.c, H3, t_demb, t_dwih, t_dbih, dtab);
    CUDA_CHECK(cudaGetLastError());
    g->launches += 2;
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    for (const float* t : {t_eemb, t_ewih, t_ebih, t_demb, t_dwih, t_dbih}) g->release((void*)t);
    m.enc_tab = etab;
    m.dec_tab = dtab;
    if (g->smem > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(g2p_words_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem));
}

static void g2p_predict(stts_g2p* g, int32_t n_words, const char* letters, const int32_t* offsets, int32_t* phones, int32_t* n_phones,
                        float* enc_hidden, float* first_logits) {
    if (n_words <= 0 || !letters || !offsets || !phones || !n_phones) throw ArgError("g2p: empty batch or null argument");
    if (offsets[0] != 0) throw ArgError("g2p: offsets[0] must be 0");
    for (int32_t w = 0; w < n_words; ++w)
        if (offsets[w + 1] <= offsets[w] || offsets[w + 1] - offsets[w] > 4096) throw ArgError("g2p: word " + std::to_string(w) + " is empty or longer than 4096 bytes");
    const int64_t nl = offsets[n_words];
    CUDA_CHECK(cudaSetDevice(g->device));
    if (nl > g->capLetters || n_words > g->capWords) {
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        for (void* p : {(void*)g->d_letters, (void*)g->d_offsets, (void*)g->d_order, (void*)g->d_preds, (void*)g->d_npreds, (void*)g->d_hidden,
                        (void*)g->d_logits})
            g->release(p);
        g->d_letters = nullptr; g->d_offsets = g->d_order = g->d_preds = g->d_npreds = nullptr; g->d_hidden = g->d_logits = nullptr;
        g->capLetters = g->capWords = 0;
        const int64_t cl = std::max<int64_t>(nl, 4096), cw = std::max<int64_t>(n_words, 256);
        g->d_letters = g->dalloc<uint8_t>(cl);
        g->d_offsets = g->dalloc<int32_t>(cw + 1);
        g->d_order = g->dalloc<int32_t>(cw);
        g->d_preds = g->dalloc<int32_t>(cw * G2P_MAX_STEPS);
        g->d_npreds = g->dalloc<int32_t>(cw);
        g->d_hidden = g->dalloc<float>(cw * g->m.H);
        g->d_logits = g->dalloc<float>(cw * g->m.Vout);
        g->capLetters = cl; g->capWords = cw;
    }
    std::vector<int32_t> order(n_words);
    for (int32_t w = 0; w < n_words; ++w) order[w] = w;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b]; });
    CUDA_CHECK(cudaMemcpyAsync(g->d_letters, letters, nl, cudaMemcpyHostToDevice, g->stream));
    CUDA_CHECK(cudaMemcpyAsync(g->d_offsets, offsets, (size_t)(n_words + 1) * 4, cudaMemcpyHostToDevice, g->stream));
    CUDA_CHECK(cudaMemcpyAsync(g->d_order, order.data(), (size_t)n_words * 4, cudaMemcpyHostToDevice, g->stream));
    CUDA_CHECK(cudaMemsetAsync(g->d_preds, 0, (size_t)n_words * G2P_MAX_STEPS * 4, g->stream));
    G2pBatch b;
    b.n_words = n_words; b.letters = g->d_letters; b.offsets = g->d_offsets; b.order = g->d_order;
    b.preds = g->d_preds; b.npreds = g->d_npreds;
    b.enc_hidden = enc_hidden ? g->d_hidden : nullptr;
    b.first_logits = first_logits ? g->d_logits : nullptr;
    const int ctas = (n_words + G2P_WPC - 1) / G2P_WPC;
    g2p_words_kernel<<<ctas, g->threads, g->smem, g->stream>>>(g->m, b);
    CUDA_CHECK(cudaGetLastError());
    g->launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(phones, g->d_preds, (size_t)n_words * G2P_MAX_STEPS * 4, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaMemcpyAsync(n_phones, g->d_npreds, (size_t)n_words * 4, cudaMemcpyDeviceToHost, g->stream));
    if (enc_hidden) CUDA_CHECK(cudaMemcpyAsync(enc_hidden, g->d_hidden, (size_t)n_words * g->m.H * 4, cudaMemcpyDeviceToHost, g->stream));
    if (first_logits) CUDA_CHECK(cudaMemcpyAsync(first_logits, g->d_logits, (size_t)n_words * g->m.Vout * 4, cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));   // `order` and the caller's buffers must outlive the copies
}

}  // namespace stts

extern "C" {

int stts_g2p_create(const float* gru_section, int64_t n_floats, int device, stts_g2p** out, int64_t* consumed_floats) {
    if (out) *out = nullptr;
    stts_g2p* g = nullptr;
    int rc = guard([&] {
        if (!gru_section || n_floats <= 0 || !out) throw ArgError("g2p: null section or output pointer");
        int ndev = 0;
        CUDA_CHECK(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev) throw ArgError("g2p: no such CUDA device");
        CUDA_CHECK(cudaSetDevice(device));
        g = new stts_g2p();
        g->device = device;
        CUDA_CHECK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
        g2p_build(g, gru_section, n_floats, consumed_floats);
        *out = g;
    });
    if (rc != STTS_OK && g) delete g;
    return rc;
}

void stts_g2p_destroy(stts_g2p* g) { delete g; }

int32_t stts_g2p_dim(const stts_g2p* g, int32_t which) {
    if (!g) return -1;
    switch (which) {
        case 0: return g->m.H;
        case 1: return g->m.Vout;
        case 2: return g->m.Vin;
        case 3: return g->m.E;
        case 4: return stts::G2P_MAX_STEPS;
        default: return -1;
    }
}

int64_t stts_g2p_kernel_launches(const stts_g2p* g) { return g ? g->launches : -1; }

int stts_g2p_predict(stts_g2p* g, int32_t n_words, const char* letters, const int32_t* offsets, int32_t* phones, int32_t* n_phones,
                     float* enc_hidden, float* first_logits) {
    if (!g) { g_last_error = "g2p: null handle"; return STTS_E_ARG; }
    return guard([&] { g2p_predict(g, n_words, letters, offsets, phones, n_phones, enc_hidden, first_logits); });
}

}  // extern "C"
