// g2p.cuh — batched GRU grapheme-to-phoneme on the GPU: the out-of-vocabulary branch of the
// reference's English frontend, /root/reference/src/engipa/EnglishText2Id.cpp:496-540 (encoder GRU over the letters,
// greedy GRU decoder, <= 20 phones), for MANY words at once.  The reference runs it one word at a time on Eigen
// (two 768x256 GEMVs per step, ~30 steps per word); at 60 000x real time for the acoustic model that serial loop is
// what a batch of English utterances with unseen words waits for.
//
// Included at the end of engine.cu (uses its CUDA_CHECK / guard / error types).  All per-thread work is in
// g2p_phases.hpp, shared with the CPU harness of the tests; this file holds the step loop, the launch and the C ABI.
//
// Two kernels, chosen per call (G2P_DEFAULT_KERNEL below); the input half of each cell is a table lookup in both
// (emb . W_ih^T + b_ih per token, built once at create time), so a step is one GEMV, not two:
//  * g2p_cluster_kernel — a thread-block cluster of 8 CTAs keeps its slices of both W_hh matrices and of fc_w in shared
//    memory for the whole launch (229 KB per CTA); the CTAs exchange hidden-state and logits slices by writing them into
//    every peer's shared memory (DSMEM) before a cluster barrier.  8 words per cluster; bound by the per-step latency
//    (shared-memory GEMV slice + cluster.sync).  Small and medium batches.
//  * g2p_words_kernel — one CTA per G2P_WPC = 4 words (sorted by length so a CTA's words finish together), 3H threads
//    (768 for the shipped model); per step every thread streams its column of W_hh (H floats, coalesced across the CTA,
//    L2-resident: 786 KB per matrix) against the four hidden states held in shared memory — one weight load feeds four
//    FMAs.  Bound by the bytes in flight per SM on the L2 -> SM stream; 296 co-resident CTAs: large batches.
#pragma once
#include <cooperative_groups.h>
#define STTS_HD __host__ __device__
#include "g2p_phases.hpp"

namespace stts {

// Kernel choice of a handle (environment STTS_G2P_KERNEL overrides the default): 0 = streaming kernel only (g2p_words_kernel),
// 1 = cluster-resident kernel only, 2 = per call: the cluster-resident kernel up to G2P_AUTO_MAX_WORDS words, the streaming kernel
// above.  Measured on B200 with the shipped model (profiles/r2_final_g2p_bench.json): 1 / 64 / 144 / 592 words 0.19 / 0.27 / 0.39 /
// 1.33 ms on the cluster kernel against 1.05 / 1.40 / 1.44 / 1.79 ms streaming; at 4096 words the streaming kernel's 296 co-resident
// CTAs win (6.2 vs 6.9 ms: 15 clusters walk 512 word groups).
constexpr int G2P_DEFAULT_KERNEL = 2;
constexpr int G2P_AUTO_MAX_WORDS = 2048;

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


// Cluster-resident variant: see g2p_phases.hpp.  Launched with a cluster dimension of G2P_CL = 8; gridDim.x / 8 clusters walk
// the word groups (G2P_WG = 8 words each, sorted by length) round-robin, so W_hh / fc_w are read from L2 once per launch.
__global__ void __launch_bounds__(256, 1) g2p_cluster_kernel(G2pDev m, G2pBatch b, int ngroups) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / G2P_CL, ncl = gridDim.x / G2P_CL;
    extern __shared__ __align__(16) float g2p_sm[];
    const G2pClDims d = g2p_cl_dims(m.H, m.Vout);
    const int H = d.H, HW = d.H * G2P_WG;
    float* We = g2p_sm;                                  // [H][R]  encoder W_hh slice
    float* Wd = We + H * d.R;                            // [H][R]  decoder W_hh slice
    float* Fw = Wd + H * d.R;                            // [H][VS] fc_w slice
    float* hb = Fw + H * d.VS;                           // [2][H][WG]
    float* gs = hb + 2 * HW;                             // [R][WG]
    float* lg = gs + d.R * G2P_WG;                       // [VS * CL][WG]
    int* tok = (int*)(lg + d.VS * G2P_CL * G2P_WG);      // [WG] (every CTA keeps the same copy)
    int* npred = tok + G2P_WG;
    int* wlen = npred + G2P_WG;
    int* woff = wlen + G2P_WG;
    int* widx = woff + G2P_WG;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int off_h = (int)(hb - g2p_sm), off_lg = (int)(lg - g2p_sm);
    float* peer[G2P_CL];                                 // every CTA's shared-memory block, as seen from here (DSMEM)
#pragma unroll
    for (int rk = 0; rk < G2P_CL; ++rk) peer[rk] = cluster.map_shared_rank(g2p_sm, rk);

    for (int i = tid; i < H * d.R; i += nt) {
        g2p_cl_load_w(i, d, rank, m.enc_whh, We);
        g2p_cl_load_w(i, d, rank, m.dec_whh, Wd);
    }
    for (int i = tid; i < H * d.VS; i += nt) g2p_cl_load_fc(i, d, rank, m.fcw, Fw);
    __syncthreads();

    for (int g = cid; g < ngroups; g += ncl) {
        for (int i = tid; i < HW; i += nt) hb[i] = 0.f;                 // h0 = 0 in buffer 0
        if (tid < G2P_WG) {
            const int s = g * G2P_WG + tid;
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
        for (int w = 0; w < G2P_WG; ++w) maxlen = max(maxlen, wlen[w]);
        cluster.sync();            // no peer writes into this CTA's buffers before they are initialised
        int cur = 0;

        for (int t = 0; t <= maxlen; ++t) {                              // encoder
            if (tid < G2P_WG) {
                const int L = wlen[tid];
                tok[tid] = (L < 0 || t > L) ? -1 : (t == L ? G2P_EOS_IN : g2p_letter_id(b.letters[woff[tid] + t]));
            }
            __syncthreads();
            for (int o = tid; o < 2 * d.R; o += nt) g2p_cl_gates_phase(o, d, rank, We, m.enc_bhh, m.enc_tab, tok, hb + cur * HW, gs);
            __syncthreads();
            {
                float* hn[G2P_CL];
#pragma unroll
                for (int rk = 0; rk < G2P_CL; ++rk) hn[rk] = peer[rk] + off_h + (cur ^ 1) * HW;
                for (int i = tid; i < d.HS * G2P_WG; i += nt) g2p_cl_update_phase(i, d, rank, m.enc_tab, tok, gs, hb + cur * HW, hn);
            }
            cluster.sync();        // every slice of the new state has landed everywhere
            cur ^= 1;
        }
        if (rank == 0 && b.enc_hidden)
            for (int i = tid; i < HW; i += nt) {
                const int j = i / G2P_WG, w = i - j * G2P_WG;
                if (wlen[w] >= 0) b.enc_hidden[(int64_t)widx[w] * H + j] = hb[cur * HW + i];
            }

        if (tid < G2P_WG) tok[tid] = wlen[tid] < 0 ? -1 : G2P_BOS_OUT;   // greedy decoder
        __syncthreads();
        for (int step = 0; step < G2P_MAX_STEPS; ++step) {
            bool any = false;
#pragma unroll
            for (int w = 0; w < G2P_WG; ++w) any |= tok[w] >= 0;       // identical in every CTA of the cluster: uniform exit
            if (!any) break;
            for (int o = tid; o < 2 * d.R; o += nt) g2p_cl_gates_phase(o, d, rank, Wd, m.dec_bhh, m.dec_tab, tok, hb + cur * HW, gs);
            __syncthreads();
            {
                float* hn[G2P_CL];
#pragma unroll
                for (int rk = 0; rk < G2P_CL; ++rk) hn[rk] = peer[rk] + off_h + (cur ^ 1) * HW;
                for (int i = tid; i < d.HS * G2P_WG; i += nt) g2p_cl_update_phase(i, d, rank, m.dec_tab, tok, gs, hb + cur * HW, hn);
            }
            cluster.sync();
            cur ^= 1;
            {
                float* lgs[G2P_CL];
#pragma unroll
                for (int rk = 0; rk < G2P_CL; ++rk) lgs[rk] = peer[rk] + off_lg;
                for (int i = tid; i < d.VS * G2P_WG; i += nt) g2p_cl_logits_phase(i, d, rank, Fw, m.fcb, tok, hb + cur * HW, lgs);
            }
            cluster.sync();        // all logits slices are in every CTA
            if (rank == 0 && step == 0 && b.first_logits)
                for (int i = tid; i < d.V * G2P_WG; i += nt) {
                    const int c = i / G2P_WG, w = i - c * G2P_WG;
                    if (wlen[w] >= 0) b.first_logits[(int64_t)widx[w] * d.V + c] = lg[i];
                }
            if (tid < G2P_WG && tok[tid] >= 0)
                g2p_cl_pick_phase(tid, d.V, lg, tok, npred, rank == 0 ? b.preds + (int64_t)widx[tid] * G2P_MAX_STEPS : nullptr);
            __syncthreads();
        }
        if (rank == 0 && tid < G2P_WG && wlen[tid] >= 0) b.npreds[widx[tid]] = npred[tid];
        __syncthreads();
    }
    cluster.sync();                // no CTA leaves while a peer may still write into its shared memory
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
    int kernel = 0;            // 0 = streaming kernel (W_hh from L2 every step), 1 = cluster-resident kernel, 2 = chosen per call
    int last_kernel = -1;      // kernel of the last predict (0 / 1)
    size_t cl_smem = 0;
    int cl_clusters = 0;       // co-resident clusters of 8 CTAs the device admits (cudaOccupancyMaxActiveClusters)

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
    // cluster-resident kernel: eligible when the slices of both W_hh, of fc_w and the state fit one CTA's shared memory
    const char* env = getenv("STTS_G2P_KERNEL");
    const int want = env ? atoi(env) : G2P_DEFAULT_KERNEL;
    g->kernel = 0;
    if ((want == 1 || want == 2) && H % G2P_CL == 0 && H % 8 == 0) {
        const G2pClDims d = g2p_cl_dims(H, fcw.r);
        const size_t bytes = (size_t)g2p_cl_smem_floats(d) * 4 + 5 * G2P_WG * 4;
        int maxOptin = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, g->device));
        if (bytes <= (size_t)maxOptin) {
            CUDA_CHECK(cudaFuncSetAttribute(g2p_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(G2P_CL);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = bytes;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = G2P_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, g2p_cluster_kernel, &cfg) == cudaSuccess && ncl > 0) {
                g->kernel = want; g->cl_smem = bytes; g->cl_clusters = ncl;
            } else {
                cudaGetLastError();    // not placeable on this device: the streaming kernel stays
            }
        }
    }
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
    const bool use_cluster = g->kernel == 1 || (g->kernel == 2 && n_words <= G2P_AUTO_MAX_WORDS);
    g->last_kernel = use_cluster ? 1 : 0;
    if (use_cluster) {
        const int ngroups = (n_words + G2P_WG - 1) / G2P_WG;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(G2P_CL * std::min(ngroups, g->cl_clusters));
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = g->cl_smem;
        cfg.stream = g->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = G2P_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, g2p_cluster_kernel, g->m, b, ngroups));
    } else {
        const int ctas = (n_words + G2P_WPC - 1) / G2P_WPC;
        g2p_words_kernel<<<ctas, g->threads, g->smem, g->stream>>>(g->m, b);
    }
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
        case 5: return g->kernel;
        case 6: return g->cl_clusters;
        case 7: return g->last_kernel;
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
