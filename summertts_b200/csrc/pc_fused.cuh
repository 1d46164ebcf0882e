// pc_fused.cuh — wide-channel Conv1d on planes with the epilogue staged through shared memory (sm_100a, tcgen05 + TMEM + TMA):
// the two convs of a WaveNet layer (WN::forward, src/modules/WN.cpp:100-149; nn_conv1d.cpp:118-199)
//
//   EPI_GATE   acts = tanh(a[:, :H] + g) * sigmoid(a[:, H:] + g),  a = conv_k5(h)              (in_layer + fused_add_tanh_sigmoid_multiply,
//                                                                                               WN.cpp:85-98,118-126)
//   EPI_RS     rs = conv_1x1(acts);  h += rs[:, :H];  skip (+)= rs[:, H:]   (last layer: skip += rs)   (WN.cpp:128-146)
//
// Like rb_fused.cuh, every tensor is split-fp16 PLANES in HBM (8 x = hi + lo, [C/8][padded row][8], hi then lo) and no epilogue
// thread touches global memory:
//   * a work item = (128-row tile of one utterance, PAIR of 64-column output chunks).  The activation tile [2 Cin/8][XR][16 B]
//     arrives by one TMA box and feeds both chunks of the pair; each chunk owns a TMEM slot: `main` (128 columns: hi*hi | hi*lo of
//     one merged N = 128 MMA per K-step) + `corr` (64 columns: lo*hi).  The two slots alternate per promotion unit (8 K-steps = 2
//     weight stages of 64 input channels), so a drain (tcgen05.ld) overlaps the other slot's MMAs.
//   * weights stream from L2 through a bulk-copy ring in 16 KB stages [64 ch / 8][hi rows 64 | lo rows 64][8 halves].
//   * EPI_RS reads-modifies-writes the residual streams IN SHARED MEMORY: the producer TMA-loads the chunk's h / skip planes tile
//     into the staging buffer, the epilogue adds the conv result in place (x8 domain), bulk async stores write it back.
//   * EPI_GATE stages 32 acts channels per chunk and bulk-stores them.
// Arithmetic modes as in rb_fused.cuh (0: merged split-fp16 + promotion, fp32-accurate; 1: one fp16 MMA per K-step).
#pragma once
#include "rb_fused.cuh"

namespace stts {

constexpr int PC_THREADS = 672;     // warps 0-15: epilogue sets (slot x column half x lane quarter); 16-19: issuers (kind x slot); 20: producer
constexpr int PC_NCH = 64;          // output columns per chunk
constexpr int PC_KC = 64;           // input channels per weight stage
constexpr int PC_STAGE = PC_KC * 2 * PC_NCH * 2;    // 16 KB
constexpr int PC_MAX_RING = 8;
enum { PC_EPI_GATE = 0, PC_EPI_RS = 1 };

struct PcWeights {
    __half* packed = nullptr;       // [chunk][tap][kc][8 groups][hi 64 | lo 64 rows][8]
    float inv_scale = 1.f;
    int k = 0, dil = 1, pad = 0, Cin = 0, Cout = 0, nchunks = 0, nst = 0;   // nst = k * Cin / 64 stages per chunk
    const float* bias = nullptr;
    bool ok = false;
};

struct PcP {
    Seg seg;
    const int2* tiles; int ntiles;   // 128-row tiles (rb_tiles_kernel with ov = 128)
    int npairs, work_items;          // chunk pairs; ntiles * npairs
    const __half* w; const float* bias; const float* gvec; int ldg;
    float isc;
    int k, dil, padl, xr, G, nst, nchunks, usteps, mode, epi, nb;
    Planes out0, out1;               // GATE: out0 = acts.  RS: out0 = h (chunks < split), out1 = skip (chunks >= split)
    int split;                       // RS: first chunk that belongs to out1
    int acc0, acc1;                  // RS: add the stream's previous value (0: store the conv result alone)
    unsigned int* flags;
    int ncta;                        // CTAs per cluster (1 or 2): with 2, each CTA fetches HALF of every weight stage and multicasts it to both
                                     // (the two CTAs work on the same chunk pair of two different row tiles), halving the L2 -> SM weight traffic
    int nclu_items;                  // cluster work items: npairs * ceil(ntiles / ncta)
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// element (row r, group g) of a staging tile [planes][groups][128 rows][16 B]
__device__ __forceinline__ uint8_t* pc_cell(uint8_t* stg, int groups, int plane, int g, int r) {
    return stg + ((size_t)(plane * groups + g) * 128 + r) * 16;
}

template <int KIND>
__device__ __forceinline__ void pc_issuer(const PcP& p, const int slot, const uint32_t a_s0, const uint32_t w_s, const uint32_t tmem,
                                          uint64_t* a_full, uint64_t* a_empty, uint64_t* m_full, uint64_t* m_empty, uint64_t* c_full,
                                          uint64_t* c_empty, uint64_t* b_full, uint64_t* b_empty) {
    const int mode = p.mode;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(((mode || KIND == 1) ? PC_NCH : 2 * PC_NCH) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t b_lbo = 2 * PC_NCH * 16;
    constexpr uint32_t b_k16 = (2 * b_lbo) >> 4;
    const uint64_t b_desc0 = ((uint64_t)((b_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46) | (uint64_t)((w_s & 0x3FFFFu) >> 4);
    const uint32_t a_lbo = (uint32_t)p.xr * 16;
    const uint64_t a_bits = ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    const uint32_t a_k16 = (2 * a_lbo) >> 4;
    const uint32_t a_s = a_s0 + ((mode == 0 && KIND == 1) ? (uint32_t)p.G * a_lbo : 0u);        // mode 0 corr reads the lo plane
    const uint32_t d_t = tmem + (uint32_t)slot * 192 + (KIND == 1 ? 128u : 0u);
    uint64_t* full_bar = KIND == 0 ? &m_full[slot] : &c_full[slot];
    uint64_t* empty_bar = KIND == 0 ? &m_empty[slot] : &c_empty[slot];
    const int SPU = mode ? (1 << 20) : max(1, p.usteps / 4);      // weight stages per promotion unit (4 K-steps per stage)
    const int kcs = p.G / 8;                                        // 64-channel blocks per tap
    uint32_t af_par = 0, e_par = 1;
    int bs = 0; uint32_t bph = 0;
    const int W = p.nclu_items, step = gridDim.x / p.ncta;
    const int rank = p.ncta == 2 ? (int)cluster_rank() : 0;
    const uint16_t mc_mask = (uint16_t)((1u << p.ncta) - 1u);
    for (int w = blockIdx.x / p.ncta; w < W; w += step) {
        const int pair = w % p.npairs, ti = (w / p.npairs) * p.ncta + rank;
        const bool active = pair * 2 + slot < p.nchunks && ti < p.ntiles;
        if (ti >= p.ntiles) {                      // a cluster's odd tail: no tile for this CTA, only the weight-stage protocol runs
            for (int s = 0; s < 2 * p.nst; ++s) {
                const int slot_r = bs; const uint32_t ph_r = bph;
                if (++bs == p.nb) { bs = 0; bph ^= 1; }
                if ((s & 1) != slot) continue;
                mbar_wait(&b_full[slot_r], ph_r);
                if (p.ncta == 2) tc_commit_mc(&b_empty[slot_r], mc_mask); else tc_commit(&b_empty[slot_r]);
            }
            continue;
        }
        mbar_wait(a_full, af_par); af_par ^= 1;
        tc_fence_after();
        const int nst = p.nst;
        const int NU = (mode || KIND == 1) ? 1 : (nst + SPU - 1) / SPU;
        int s_done = 0;
#pragma unroll 1
        for (int un = 0; un < NU; ++un) {
            const int s1 = NU == 1 ? nst : min(nst, s_done + SPU);
            if (active) { mbar_wait(empty_bar, e_par); e_par ^= 1; tc_fence_after(); }
            uint32_t acc = 0u;
#pragma unroll 1
            for (int s = s_done; s < s1; ++s) {
                // ring entries alternate slot 0 / slot 1 of the same stage index; every issuer walks all of them
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    const int slot_r = bs; const uint32_t ph_r = bph;
                    if (++bs == p.nb) { bs = 0; bph ^= 1; }
                    if (q != slot) continue;
                    if (mode && ((s & 1) != KIND)) {          // not this issuer's stage: just release it
                        mbar_wait(&b_full[slot_r], ph_r);
                        if (p.ncta == 2) tc_commit_mc(&b_empty[slot_r], mc_mask); else tc_commit(&b_empty[slot_r]);
                        continue;
                    }
                    mbar_wait(&b_full[slot_r], ph_r);
                    tc_fence_after();
                    if (active) {
                        const int tap = s / kcs, kc = s - tap * kcs;
                        const uint64_t da = a_bits | (uint64_t)(((a_s + (uint32_t)(tap * p.dil) * 16 + (uint32_t)(kc * 8) * a_lbo) & 0x3FFFFu) >> 4);
                        const uint64_t db = b_desc0 + (uint32_t)slot_r * (uint32_t)(PC_STAGE >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_t, da + (uint32_t)(ks * a_k16), db + (uint32_t)(ks * b_k16), idesc, ks == 0 ? acc : 1u);
                        acc = 1u;
                    }
                    if (p.ncta == 2) tc_commit_mc(&b_empty[slot_r], mc_mask); else tc_commit(&b_empty[slot_r]);
                }
            }
            s_done = s1;
            if (active) tc_commit(full_bar);
        }
        tc_commit(a_empty);          // this issuer's reads of the activation tile have retired
    }
}

template <int EPI>
__global__ void __launch_bounds__(PC_THREADS, 1) pc_kernel(const PcP p, const __grid_constant__ CUtensorMap imap, const __grid_constant__ CUtensorMap rmap0,
                                                           const __grid_constant__ CUtensorMap rmap1) {
    constexpr int SG = EPI == PC_EPI_GATE ? 4 : 8;          // 16-byte channel groups per plane of a staging tile (32 / 64 channels)
    constexpr int STG = 2 * SG * 128 * 16;                   // bytes per slot: 16 KB / 32 KB
    extern __shared__ __align__(128) uint8_t psm[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_max_sync(0xffffffffu, tid >> 5);
    const uint32_t a_tile = (uint32_t)p.G * 2 * p.xr * 16;
    uint8_t* abuf = psm;
    uint8_t* stg = abuf + a_tile;                             // [2 slots][STG]
    uint8_t* wst = stg + 2 * STG;                             // [nb][PC_STAGE]
    uint64_t* bars = reinterpret_cast<uint64_t*>(wst + (size_t)p.nb * PC_STAGE);
    uint64_t* a_full = bars;             // [1]
    uint64_t* a_empty = bars + 1;        // [1]
    uint64_t* m_full = bars + 2;         // [2]
    uint64_t* m_empty = bars + 4;        // [2]
    uint64_t* c_full = bars + 6;         // [2]
    uint64_t* c_empty = bars + 8;        // [2]
    uint64_t* r_full = bars + 10;        // [2] residual tile landed in the slot's staging buffer (RS)
    uint64_t* s_free = bars + 12;        // [2] staging buffer free again (its bulk stores have read it)
    uint64_t* b_full = bars + 14;        // [PC_MAX_RING]
    uint64_t* b_empty = b_full + PC_MAX_RING;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + PC_MAX_RING);

    if (tid == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&m_full[i], 1); mbar_init(&m_empty[i], 8);
            mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 8);
            mbar_init(&r_full[i], 1); mbar_init(&s_free[i], 1);
        }
        for (int s = 0; s < PC_MAX_RING; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2 * p.ncta); }    // both issuers of the slot, in every CTA of the cluster
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;     // slot s: main @ s*192 (128 columns: hi*hi | hi*lo), corr @ s*192 + 128 (64 columns)
    if (p.ncta == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into them
    const int W = p.nclu_items, wstep = gridDim.x / p.ncta;
    const int crank = p.ncta == 2 ? (int)cluster_rank() : 0;

    if (warp < 16) {
        // ================= promotion + epilogue: (slot, column half hf), one row x 32 columns per thread ==================
        const int wq = warp & 3, slot = (warp >> 2) & 1, hf = warp >> 3;
        const int tl = wq * 32 + lane;
        const uint32_t tmain = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(slot * 192 + hf * 32);
        const uint32_t tcorr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(slot * 192 + 128 + hf * 32);
        uint8_t* sbuf = stg + (size_t)slot * STG;
        uint32_t mf_par = 0, cf_par = 0, rf_par = 0;
        float amax = 0.f;
        float racc[32];
        const int SPU = p.mode ? (1 << 20) : max(1, p.usteps / 4);
        const int NU = p.mode ? 1 : (p.nst + SPU - 1) / SPU;
        int item = 0;
        for (int w = blockIdx.x / p.ncta; w < W; w += wstep, ++item) {
            const int pair = w % p.npairs, ti = (w / p.npairs) * p.ncta + crank;
            const int chunk = pair * 2 + slot;
            const bool active = chunk < p.nchunks && ti < p.ntiles;
            if (!active) continue;                    // (the partner slot still runs; nothing of this slot's barriers is used)
            const RbTile it = rb_tile_at(p.seg, p.tiles, ti);
            for (int un = 0; un < NU; ++un) {
                mbar_wait_all(&m_full[slot], mf_par); mf_par ^= 1;
                tc_fence_after();
#pragma unroll
                for (int cb = 0; cb < 32; cb += 16) {
                    uint32_t v[16], x2[16];
                    tc_ld_nowait<16>(tmain + cb, v);
                    if (!p.mode) tc_ld_nowait<16>(tmain + 64 + cb, x2);
                    tc_ld_wait();
                    if (p.mode) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[cb + j] = __uint_as_float(v[j]);
                    } else if (un == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[cb + j] = __uint_as_float(v[j]) + __uint_as_float(x2[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[cb + j] += __uint_as_float(v[j]) + __uint_as_float(x2[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&m_empty[slot]);
            }
            {
                mbar_wait_all(&c_full[slot], cf_par); cf_par ^= 1;
                tc_fence_after();
#pragma unroll
                for (int cb = 0; cb < 32; cb += 16) {
                    uint32_t v[16];
                    tc_ld_nowait<16>(tcorr + cb, v);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) racc[cb + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&c_empty[slot]);
            }
            const int tr = it.t0 + tl;
            const bool valid = tr < it.len;
            const int n0 = chunk * PC_NCH + hf * 32;       // first output column of this thread
            if (EPI == PC_EPI_GATE) {
                // the previous item's bulk stores of this slot (issued by these same lanes) must have read the staging tile
                if (tl < 2 * SG && hf == 0) bulk_wait_read0();
                asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
                const float isc = p.isc;
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float a = fmaf(racc[2 * j], isc, p.bias ? __ldg(p.bias + n0 + 2 * j) : 0.f);
                    float b = fmaf(racc[2 * j + 1], isc, p.bias ? __ldg(p.bias + n0 + 2 * j + 1) : 0.f);
                    if (p.gvec) { a += __ldg(p.gvec + (size_t)it.u * p.ldg + n0 + 2 * j); b += __ldg(p.gvec + (size_t)it.u * p.ldg + n0 + 2 * j + 1); }
                    const float g = gate_ref(a, b) * TC_ASCALE;
                    amax = fmaxf(amax, valid ? fabsf(g) : 0.f);
                    o[j] = g;
                }
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    uint4 hi, lo;
                    split8_scaled(o + 8 * gg, hi, lo);
                    if (!valid) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
                    *reinterpret_cast<uint4*>(pc_cell(sbuf, SG, 0, hf * 2 + gg, tl)) = hi;
                    *reinterpret_cast<uint4*>(pc_cell(sbuf, SG, 1, hf * 2 + gg, tl)) = lo;
                }
                fence_proxy_async();
                asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
                if (tl < 2 * SG && hf == 0) {             // 8 bulk stores per chunk: (plane, group) x the valid rows
                    const int plane = tl / SG, g = tl - plane * SG;
                    const int nrows = min(128, it.len - it.t0);
                    const int Gout = p.out0.C / 8;
                    __half* gdst = p.out0.base + ((size_t)(plane * Gout + chunk * SG + g) * p.out0.rows_p + (size_t)(it.prow_u + it.t0)) * 8;
                    bulk_s2g(gdst, pc_cell(sbuf, SG, plane, g, 0), (uint32_t)nrows * 16);
                    bulk_commit();
                }
            } else {
                const bool to1 = chunk >= p.split;
                const Planes& op = to1 ? p.out1 : p.out0;
                const int oc = to1 ? chunk - p.split : chunk;          // 64-channel chunk inside the destination stream
                const bool accin = to1 ? p.acc1 != 0 : p.acc0 != 0;
                if (accin) { mbar_wait_all(&r_full[slot], rf_par); rf_par ^= 1; }
                else {                                                  // nothing was loaded: the staging tile is written from scratch
                    if (tl < 2 * SG && hf == 0) bulk_wait_read0();
                    asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
                }
                const float isc8 = p.isc * TC_ASCALE;
#pragma unroll
                for (int gg = 0; gg < 4; ++gg) {
                    uint8_t* ch = pc_cell(sbuf, SG, 0, hf * 4 + gg, tl);
                    uint8_t* cl = pc_cell(sbuf, SG, 1, hf * 4 + gg, tl);
                    uint4 xh = make_uint4(0, 0, 0, 0), xl = xh;
                    if (accin) { xh = *reinterpret_cast<const uint4*>(ch); xl = *reinterpret_cast<const uint4*>(cl); }
                    const uint32_t hh[4] = {xh.x, xh.y, xh.z, xh.w}, ll[4] = {xl.x, xl.y, xl.z, xl.w};
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hh[j]));
                        const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ll[j]));
                        const int n = n0 + 8 * gg + 2 * j;
                        const float y0 = fmaf(racc[8 * gg + 2 * j], isc8, (p.bias ? __ldg(p.bias + n) : 0.f) * TC_ASCALE) + (fh.x + fl.x);
                        const float y1 = fmaf(racc[8 * gg + 2 * j + 1], isc8, (p.bias ? __ldg(p.bias + n + 1) : 0.f) * TC_ASCALE) + (fh.y + fl.y);
                        amax = fmaxf(amax, valid ? fmaxf(fabsf(y0), fabsf(y1)) : 0.f);
                        o[2 * j] = y0; o[2 * j + 1] = y1;
                    }
                    uint4 hi, lo;
                    split8_scaled(o, hi, lo);
                    if (!valid) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
                    *reinterpret_cast<uint4*>(ch) = hi;
                    *reinterpret_cast<uint4*>(cl) = lo;
                }
                fence_proxy_async();
                asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
                if (tl < 2 * SG && hf == 0) {             // 16 bulk stores per chunk
                    const int plane = tl / SG, g = tl - plane * SG;
                    const int nrows = min(128, it.len - it.t0);
                    const int Gout = op.C / 8;
                    __half* gdst = op.base + ((size_t)(plane * Gout + oc * SG + g) * op.rows_p + (size_t)(it.prow_u + it.t0)) * 8;
                    bulk_s2g(gdst, pc_cell(sbuf, SG, plane, g, 0), (uint32_t)nrows * 16);
                    bulk_commit();
                    bulk_wait_read0();                    // the producer may refill this staging tile once every store has read it
                }
                if (wq == 0 && hf == 0) {                 // the storing warp alone waits; the other warps go on to the next item's drains
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_free[slot]);
                }
            }
        }
        if (tl < 2 * SG && hf == 0) bulk_wait_all0();
        if (amax > 65000.f && p.flags) atomicOr(p.flags, 1u);
    } else if (warp < 20) {
        if (lane == 0) {
            const int kind = (warp - 16) >> 1, islot = (warp - 16) & 1;
            if (kind == 0) pc_issuer<0>(p, islot, smem_u32(abuf), smem_u32(wst), tmem, a_full, a_empty, m_full, m_empty, c_full, c_empty, b_full, b_empty);
            else pc_issuer<1>(p, islot, smem_u32(abuf), smem_u32(wst), tmem, a_full, a_empty, m_full, m_empty, c_full, c_empty, b_full, b_empty);
        }
        __syncwarp();
    } else {
        // ================= producer: activation tiles, residual tiles (RS) and weight stages, one thread, cooperative polling ===
        if (lane == 0) {
            int xw = blockIdx.x / p.ncta; uint32_t ae_par = 1;
            int rw = blockIdx.x / p.ncta; uint32_t sf_par0 = 1, sf_par1 = 1; int rslot = 0;
            int ww = blockIdx.x / p.ncta, wstage = 0, wq2 = 0, ws_ = 0; uint32_t wph = 1;
            const uint16_t mc_mask = (uint16_t)((1u << p.ncta) - 1u);
            if (EPI != PC_EPI_RS) rw = W;
            while (xw < W || ww < W || rw < W) {
                bool progress = false;
                if (xw < W && (xw / p.npairs) * p.ncta + crank >= p.ntiles) { xw += wstep; progress = true; }      // odd tail: no tile for this CTA
                else if (xw < W && mbar_test(a_empty, ae_par)) {
                    ae_par ^= 1;
                    const RbTile it = rb_tile_at(p.seg, p.tiles, (xw / p.npairs) * p.ncta + crank);
                    const long long r0 = it.prow_u + it.t0 - p.padl;
                    mbar_expect_tx(a_full, a_tile);
                    tma_load_3d(abuf, &imap, 0, (int)r0, 0, a_full);
                    xw += wstep;
                    progress = true;
                }
                if (rw < W) {                          // residual tile of (item rw, slot rslot): hi groups, then lo groups
                    const int chunk = (rw % p.npairs) * 2 + rslot;
                    const bool to1 = chunk >= p.split;
                    const int rti = (rw / p.npairs) * p.ncta + crank;
                    const bool act = chunk < p.nchunks && rti < p.ntiles;
                    const bool need = act && (to1 ? p.acc1 != 0 : p.acc0 != 0);
                    bool adv = !act;
                    // every active chunk's epilogue frees the staging tile once (s_free); the producer consumes each of those
                    // completions, loading the stream's previous value when the chunk accumulates
                    if (act && mbar_test(&s_free[rslot], rslot ? sf_par1 : sf_par0)) {
                        if (rslot) sf_par1 ^= 1; else sf_par0 ^= 1;
                        adv = true;
                        progress = true;
                    }
                    if (act && adv && need) {
                        const RbTile it = rb_tile_at(p.seg, p.tiles, rti);
                        const Planes& op = to1 ? p.out1 : p.out0;
                        const int oc = to1 ? chunk - p.split : chunk;
                        const CUtensorMap* rm = to1 ? &rmap1 : &rmap0;
                        uint8_t* dst = stg + (size_t)rslot * STG;
                        mbar_expect_tx(&r_full[rslot], STG);
                        tma_load_3d(dst, rm, 0, (int)(it.prow_u + it.t0), oc * SG, &r_full[rslot]);
                        tma_load_3d(dst + STG / 2, rm, 0, (int)(it.prow_u + it.t0), op.C / 8 + oc * SG, &r_full[rslot]);
                    }
                    if (adv) { if (++rslot == 2) { rslot = 0; rw += wstep; } }
                }
                if (ww < W && mbar_test(&b_empty[ws_], wph)) {
                    // ring order: stage 0 of slot 0, stage 0 of slot 1, stage 1 of slot 0, ...
                    const int chunk = (ww % p.npairs) * 2 + wq2;
                    const int cc = min(chunk, p.nchunks - 1);          // an inactive slot still gets (ignored) bytes: keeps the ring walk uniform
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w) + ((size_t)cc * p.nst + wstage) * PC_STAGE;
                    mbar_expect_tx(&b_full[ws_], PC_STAGE);
                    if (p.ncta == 2) {        // this CTA fetches its half of the stage and multicasts it into both CTAs (same offsets, same barrier)
                        const uint32_t half = PC_STAGE / 2;
                        bulk_g2s_mc(wst + (size_t)ws_ * PC_STAGE + crank * half, src + crank * half, half, &b_full[ws_], mc_mask);
                    } else bulk_g2s(wst + (size_t)ws_ * PC_STAGE, src, PC_STAGE, &b_full[ws_]);
                    if (++ws_ == p.nb) { ws_ = 0; wph ^= 1; }
                    if (++wq2 == 2) { wq2 = 0; if (++wstage == p.nst) { wstage = 0; ww += wstep; } }
                    progress = true;
                }
                if (!progress) __nanosleep(32);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (p.ncta == 2) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
inline void pc_prepare_weights(PcWeights& r, const float* w /*[k][Cin][CoutW]*/, int k, int Cin, int Cout, int CoutW, int dil, int pad,
                               const float* bias_dev, std::vector<void*>& owned) {
    r.ok = false;
    if (Cin % PC_KC != 0 || Cout % PC_NCH != 0 || k < 1 || k > 16) return;
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)k * Cin * CoutW; ++i) mx = std::max(mx, std::fabs(w[i]));
    int e = 0;
    if (mx > 0.f) e = 9 - (int)std::floor(std::log2(mx));
    e = std::max(-10, std::min(e, 20));
    const float ws = std::ldexp(1.0f, e);
    const int nchunks = Cout / PC_NCH, kcs = Cin / PC_KC, nst = k * kcs;
    std::vector<__half> buf((size_t)nchunks * nst * (PC_STAGE / 2));
    for (int c = 0; c < nchunks; ++c)
        for (int tap = 0; tap < k; ++tap)
            for (int kc = 0; kc < kcs; ++kc) {
                __half* dst = buf.data() + ((size_t)c * nst + tap * kcs + kc) * (PC_STAGE / 2);
                for (int ch = 0; ch < PC_KC; ++ch)
                    for (int n = 0; n < PC_NCH; ++n) {
                        const float v = w[((size_t)tap * Cin + kc * PC_KC + ch) * CoutW + c * PC_NCH + n] * ws;
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        const size_t ih = ((size_t)(ch / 8) * 2 * PC_NCH + n) * 8 + (ch % 8);
                        dst[ih] = hi;
                        dst[ih + (size_t)PC_NCH * 8] = lo;
                    }
            }
    void* d = nullptr;
    if (cudaMalloc(&d, buf.size() * sizeof(__half)) != cudaSuccess) return;
    owned.push_back(d);
    if (cudaMemcpy(d, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return;
    r.packed = (__half*)d; r.inv_scale = std::ldexp(1.0f, -e) / TC_ASCALE; r.k = k; r.dil = dil; r.pad = pad; r.Cin = Cin; r.Cout = Cout;
    r.nchunks = nchunks; r.nst = nst; r.bias = bias_dev;
    r.ok = true;
}
inline cudaError_t pc_device_setup() {
    cudaError_t e = cudaFuncSetAttribute(pc_kernel<PC_EPI_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pc_kernel<PC_EPI_RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
struct PcPlan { size_t smem; int nb; };
inline PcPlan pc_plan(const PcWeights& w, int epi) {
    PcPlan pl;
    const size_t xr = 128 + (size_t)(w.k - 1) * w.dil;
    const size_t a_tile = (size_t)(w.Cin / 8) * 2 * xr * 16;
    const size_t stg = (size_t)2 * (epi == PC_EPI_GATE ? 16 : 32) * 1024;
    const size_t misc = (14 + 2 * PC_MAX_RING) * 8 + 64;
    const size_t budget = 225 * 1024;
    size_t room = budget > a_tile + stg + misc ? (budget - a_tile - stg - misc) / PC_STAGE : 0;
    pl.nb = (int)std::min<size_t>(room, PC_MAX_RING);
    pl.nb &= ~1;                      // even: a stage pair (slot 0, slot 1) never straddles the wrap
    pl.smem = a_tile + stg + (size_t)pl.nb * PC_STAGE + misc;
    return pl;
}
inline bool pc_eligible(const PcWeights& w, int epi) {
    if (!w.ok || w.pad > TC_GAP || 128 + (w.k - 1) * w.dil > 256 || 2 * w.pad != (w.k - 1) * w.dil) return false;
    if (2 * (w.Cin / 8) > 256) return false;
    if (w.nst < 2) return false;       // throughput mode splits the stages between the two accumulators
    return pc_plan(w, epi).nb >= 2;
}
inline bool pc_make_map(CUtensorMap* m, const Planes& pl, int box_rows, int box_groups) {
    cuuint64_t dims[3] = {8, (cuuint64_t)pl.rows_p, (cuuint64_t)(2 * (pl.C / 8))};
    cuuint64_t strides[2] = {16, (cuuint64_t)pl.rows_p * 16};
    cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_groups};
    cuuint32_t estr[3] = {1, 1, 1};
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
        return (EncodeFn)fn;
    }();
    if (!encode) return false;
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, pl.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// GATE: in = h planes, out0 = acts planes (Cout / 2 channels), gvec = per-utterance conditioning (interleaved like the weights) or null.
// RS:   in = acts planes, out0 = h planes (updated in place), out1 = skip planes; split_ch = Cout columns that go to h (0: all to skip);
//       acc0 / acc1: add the previous value of the stream.
inline int pc_launch(int epi, const PcWeights& w, const Planes& in, const Planes& out0, const Planes& out1, int split_ch, int acc0, int acc1,
                     const float* gvec, int ldg, Seg seg, const int2* tiles, int ntiles, int mode, int sms, unsigned int* flags, cudaStream_t stream) {
    if (!pc_eligible(w, epi) || in.C != w.Cin || ntiles <= 0 || !tiles) return -3;
    PcP p;
    p.seg = seg; p.tiles = tiles; p.ntiles = ntiles;
    p.npairs = (w.nchunks + 1) / 2;
    p.work_items = ntiles * p.npairs;
    p.w = w.packed; p.bias = w.bias; p.gvec = gvec; p.ldg = ldg; p.isc = w.inv_scale;
    p.k = w.k; p.dil = w.dil; p.padl = w.pad; p.xr = 128 + (w.k - 1) * w.dil; p.G = w.Cin / 8; p.nst = w.nst; p.nchunks = w.nchunks;
    static const int env_us = getenv("STTS_TC_USTEPS") ? atoi(getenv("STTS_TC_USTEPS")) : 0;
    p.usteps = env_us > 0 ? env_us : 8;
    p.mode = mode; p.epi = epi;
    const PcPlan pl = pc_plan(w, epi);
    p.nb = pl.nb;
    p.out0 = out0; p.out1 = out1; p.split = split_ch / PC_NCH; p.acc0 = acc0; p.acc1 = acc1; p.flags = flags;
    if (epi == PC_EPI_GATE) { p.split = 1 << 20; if (out0.C * 2 != w.Cout) return -3; }
    else if (split_ch % PC_NCH != 0 || (split_ch > 0 && out0.C != split_ch) || out1.C != w.Cout - split_ch) return -3;
    alignas(64) CUtensorMap imap, rmap0, rmap1;
    if (!pc_make_map(&imap, in, p.xr, 2 * p.G)) return -1;
    const Planes& r0 = (epi == PC_EPI_RS && split_ch > 0) ? out0 : out1;
    if (!pc_make_map(&rmap0, epi == PC_EPI_RS ? r0 : in, 128, 8) || !pc_make_map(&rmap1, epi == PC_EPI_RS ? out1 : in, 128, 8)) return -1;
    static const int env_clu = getenv("STTS_PC_CLUSTER") ? atoi(getenv("STTS_PC_CLUSTER")) : 1;   // 2: weight multicast in CTA pairs (works; measured no gain: the ring depth, not L2 bandwidth, bounds these kernels)
    p.ncta = (env_clu == 2 && ntiles >= 2) ? 2 : 1;
    p.nclu_items = p.npairs * ((ntiles + p.ncta - 1) / p.ncta);
    int ctas = std::min(sms, p.nclu_items * p.ncta);
    ctas -= ctas % p.ncta;
    static const int env_verbose = getenv("STTS_TC_VERBOSE") ? atoi(getenv("STTS_TC_VERBOSE")) : 0;
    if (env_verbose > 0) {
        static int left = 8;
        if (left > 0) { --left; fprintf(stderr, "pc_conv: epi=%d Cin=%d Cout=%d k=%d mode=%d items=%d ctas=%d smem=%zu nb=%d nst=%d\n", epi, w.Cin, w.Cout, w.k, mode, p.work_items, ctas, pl.smem, pl.nb, w.nst); }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(PC_THREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (p.ncta == 2) {
        // persistent CTAs walk the items with a static stride: every cluster must be resident at once, or the late ones
        // serialise behind a whole pass of the others.  GPCs with an odd SM count leave some SMs without a partner.
        int maxc = 0;
        const cudaError_t oe = epi == PC_EPI_GATE ? cudaOccupancyMaxActiveClusters(&maxc, pc_kernel<PC_EPI_GATE>, &cfg)
                                                 : cudaOccupancyMaxActiveClusters(&maxc, pc_kernel<PC_EPI_RS>, &cfg);
        if (oe == cudaSuccess && maxc > 0 && 2 * maxc < ctas) { ctas = 2 * maxc; cfg.gridDim = dim3(ctas); }
    }
    if (env_verbose > 0) {
        static int left2 = 4;
        if (left2 > 0) { --left2; fprintf(stderr, "pc_conv: cluster %d, grid %d, cluster items %d\n", p.ncta, ctas, p.nclu_items); }
    }
    cudaError_t le;
    if (epi == PC_EPI_GATE) le = cudaLaunchKernelEx(&cfg, pc_kernel<PC_EPI_GATE>, p, imap, rmap0, rmap1);
    else le = cudaLaunchKernelEx(&cfg, pc_kernel<PC_EPI_RS>, p, imap, rmap0, rmap1);
    return le == cudaSuccess ? 1 : -5;
}

}  // namespace stts
