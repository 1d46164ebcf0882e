// rb_fused.cuh — one ResBlock1 "pair" fused in a persistent tcgen05 kernel (sm_100a):
//
//     x' = x + conv2( leaky_0.1( conv1_dil( leaky_0.1(x) ) ) )          (ResBlock1::forward, src/modules/ResBlock1.cpp:55-69,
//                                                                         one iteration of its loop; nn_conv1d.cpp:118-199 twice)
//
// The residual stream never exists as fp32 in HBM.  It lives as split-fp16 planes of act(x) (act = leaky 0.1 between
// pairs, identity after the last pair of a ResBlock1): plane[c/8][padded row][c%8], hi then lo, 8*act(x) = hi + lo
// (22 significant bits; see conv_tc.cuh).  Per 256-row super-tile a CTA
//   1. TMA-loads the x planes for two 128-row M-tiles (own halos) straight into the UMMA no-swizzle K-major layout,
//   2. conv1: taps = descriptor row shifts of that tile (dilation d), accumulators in TMEM,
//   3. epilogue 1: bias + leaky -> split-fp16 -> SHARED MEMORY tile T1 (256 + k-1 rows), zero outside the utterance,
//   4. conv2: A operand = T1 (dilation 1),
//   5. epilogue 2: bias + residual (recovered from the x tile still in shared memory: x = unleaky((hi+lo)/8)) -> act ->
//      split-fp16 -> shared memory -> bulk async stores (one per 16-byte channel group and plane) of the valid rows.
// HBM traffic per pair: planes in (4 B/element + halo) + planes out (4 B/element); no epilogue thread touches global memory.
//
// Arithmetic modes
//   mode 0 (default, fp32-accurate): merged split-fp16: per 16-channel K-step one N = 2C MMA  A_hi x [W_hi | W_lo]  into `main`
//       and one N = C MMA  A_lo x W_hi  into `corr` (dropped lo*lo term: 2^-22).  `main` is promoted to fp32 registers every
//       `usteps` K-steps (the tensor core's fp32 accumulator truncates per MMA; conv_tc.cuh).  The two M-tiles alternate per
//       promotion unit, so a drain (tcgen05.ld) always overlaps the other tile's MMAs and one `main` per M-tile suffices.
//   mode 1 (throughput): one N = C MMA per K-step on the hi planes only (fp16 operands, fp32 accumulate); the residual stream
//       keeps its hi+lo precision.  The two issuer warps take one M-tile each.
// Measured MMA cost model on B200 (tools/mma_bench2/3.cu): an M=128,K=16 kind::f16 MMA retires in max(N/2, 32 + N/4) cycles,
// independent of the number of issuing warps; a single issuing thread cannot go below ~50 cycles per MMA.
#pragma once
#include "conv_tc.cuh"

namespace stts {

constexpr int RB_THREADS = 384;     // warps 0-3 / 4-7: epilogue sets of M-tile 0 / 1; 8, 9: MMA issuers; 10: x-tile TMA; 11: weights
constexpr int RB_MAX_STAGES = 32;   // weight stages (taps of conv1 + conv2) when resident; ring depth otherwise

struct RbWeights {                  // one conv of the pair, merged split-fp16 stages [tap][C/8][hi rows C | lo rows C][8]
    __half* packed = nullptr;
    float inv_scale = 1.f;
    int k = 0, dil = 1, pad = 0;
    const float* bias = nullptr;
    bool ok = false;
};

struct RbP {
    Seg seg;
    int gx, work_items;              // super-tiles of the longest utterance; gx * utterances
    const __half* w1; const __half* w2;
    const float* b1; const float* b2;
    float isc1, isc2;
    int k1, d1, pad1, k2, pad2;
    int xr1, xr2, ov;                // x-tile rows per M-tile (128 + (k1-1) d1), T1 rows (256 + k2 - 1), valid output rows (256 - (k2-1))
    int usteps;
    int mode;                        // 0 accurate (merged split-fp16), 1 throughput (single fp16 MMA)
    int abufs;                       // x-tile buffers (1 or 2)
    int resident, nb;                // weights resident (nb = k1 + k2 stages) or ring of nb stages
    float in_slope;                  // input planes hold leaky(x): residual x = v < 0 ? v / in_slope : v   (0: identity)
    int out_act; float out_slope;
    int tmem_cols;
    Planes outp;                     // destination planes (written with 1-D bulk stores of the exact valid rows)
    unsigned int* flags;             // bit 0: an activation exceeded the fp16 range of the split (|8 x| > 65504)
    long long* trace;                // optional clock64 timeline of CTA 0 (STTS_TC_TRACE_BUILD + STTS_RB_TRACE): [6 roles][1024]
};

#ifdef STTS_TC_TRACE_BUILD
#define RB_TS(role, idx) do { if (rtr && (idx) < 1024) rtr[(role) * 1024 + (idx)] = clock64(); } while (0)
#else
#define RB_TS(role, idx) do { (void)rtr; } while (0)
#endif

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// lane 0 polls, the warp re-converges, then every lane observes the (already complete) phase itself
__device__ __forceinline__ void mbar_wait_all(uint64_t* b, uint32_t parity) {
    mbar_wait_warp(b, parity);
    mbar_wait(b, parity);
}

struct RbTile { int w, W, step, gx; int u, x, seg0, len, t0; long long prow_u; };
__device__ __forceinline__ bool rb_next(RbTile& it, const Seg& seg, int ov) {
    while (it.w < it.W) {
        const int w = it.w;
        it.w += it.step;
        const int u = w / it.gx, x = w - u * it.gx;
        const int len = seg_len(seg, u);
        if (x * ov >= len) continue;
        it.u = u; it.x = x; it.len = len; it.t0 = x * ov;
        it.seg0 = seg_start(seg, u);
        it.prow_u = planes_row(seg, u);
        return true;
    }
    return false;
}

template <int C>
__global__ void __launch_bounds__(RB_THREADS, 1) rb_pair_kernel(const RbP p, const __grid_constant__ CUtensorMap imap) {
    constexpr int G = C / 8;          // 16-byte channel groups per plane
    constexpr int KS = C / 16;        // K = 16 MMA steps per tap
    extern __shared__ __align__(128) uint8_t rsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int XR1 = p.xr1, XR2 = p.xr2, OV = p.ov;
    const uint32_t a_tile = (uint32_t)C * XR1 * 4;           // bytes of one M-tile's x planes: [2G][XR1][16 B]
    const uint32_t t1_bytes = (uint32_t)C * XR2 * 4;         // [2G][XR2][16 B]   (aliased by the output tile [2G][OV][16 B])
    const uint32_t stage = (uint32_t)4 * C * C;              // one tap: [G][2C rows][16 B]
    uint8_t* abuf = rsm;                                      // [abufs][2 M-tiles][a_tile]
    uint8_t* t1 = abuf + (size_t)p.abufs * 2 * a_tile;
    uint8_t* wst = t1 + t1_bytes;                             // [nb][stage]
    uint64_t* bars = reinterpret_cast<uint64_t*>(wst + (size_t)p.nb * stage);
    uint64_t* a_full = bars;             // [2]
    uint64_t* a_empty = bars + 2;        // [2]
    uint64_t* m_full = bars + 4;         // [2] per M-tile
    uint64_t* m_empty = bars + 6;        // [2]
    uint64_t* c_full = bars + 8;         // [2]
    uint64_t* c_empty = bars + 10;       // [2]
    uint64_t* t1_full = bars + 12;       // [1]
    uint64_t* b_full = bars + 13;        // [RB_MAX_STAGES]
    uint64_t* b_empty = b_full + RB_MAX_STAGES;   // [RB_MAX_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + RB_MAX_STAGES);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);   // [2][C]

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2);
            mbar_init(&m_full[i], 1); mbar_init(&m_empty[i], 128);
            mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 128);
        }
        mbar_init(t1_full, 256);
        for (int s = 0; s < RB_MAX_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * C) sbias[tid] = tid < C ? (p.b1 ? __ldg(p.b1 + tid) : 0.f) : (p.b2 ? __ldg(p.b2 + tid - C) : 0.f);
    // rows >= 256 of T1 are never produced by conv1; they only feed output rows >= OV, which are discarded, but must not
    // hold NaN patterns that a later kernel version might read: zero the whole buffer once
    for (uint32_t i = tid; i < t1_bytes / 16; i += RB_THREADS) reinterpret_cast<uint4*>(t1)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;      // M-tile mt: main @ mt*3C (2C columns), corr @ mt*3C + 2C (C columns)

    RbTile it;
    it.w = blockIdx.x; it.W = p.work_items; it.step = gridDim.x; it.gx = p.gx;
    long long* rtr = (p.trace && blockIdx.x == 0 && (threadIdx.x & 127) == 0) ? p.trace : nullptr;   // tid 0, 128 (sets), 256 (issuer 0), ... see roles below
    const int UPT = p.mode ? 64 : max(1, p.usteps / KS);     // taps per promotion unit

    if (warp < 8) {
        // ================= promotion + epilogues of M-tile `mt` ======================================
        const int mt = warp >> 2, wq = warp & 3, tl = tid & 127;
        const uint32_t tmain = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * 3 * C);
        const uint32_t tcorr = tmain + 2 * C;
        uint32_t mf_par = 0, cf_par = 0;
        bool ovf = false;
        float racc[C];
        for (int tile = 0; rb_next(it, p.seg, OV); ++tile) {
            const int buf = p.abufs == 2 ? (tile & 1) : 0;
            for (int ph = 0; ph < 2; ++ph) {
                const int k = ph ? p.k2 : p.k1;
                const int NU = p.mode ? 1 : (k + UPT - 1) / UPT;
                for (int un = 0; un < NU; ++un) {
                    mbar_wait_all(&m_full[mt], mf_par); mf_par ^= 1;
                    tc_fence_after();
                    if (un == 0) RB_TS(2 + mt, tile * 8 + ph * 3);
                    if (p.mode) {
#pragma unroll
                        for (int cb = 0; cb < C; cb += 16) tc_ld16(tmain + cb, racc + cb);
                    } else {
#pragma unroll
                        for (int cb = 0; cb < C; cb += 16) {
                            float v[16], x2[16];
                            tc_ld16(tmain + cb, v);
                            tc_ld16(tmain + C + cb, x2);
                            if (un == 0) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) racc[cb + j] = v[j] + x2[j];
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) racc[cb + j] += v[j] + x2[j];
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&m_empty[mt]);
                }
                if (!p.mode) {
                    mbar_wait_all(&c_full[mt], cf_par); cf_par ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int cb = 0; cb < C; cb += 16) {
                        float v[16];
                        tc_ld16(tcorr + cb, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[cb + j] += v[j];
                    }
                    tc_fence_before();
                    mbar_arrive(&c_empty[mt]);
                }
                RB_TS(2 + mt, tile * 8 + ph * 3 + 1);
                const int i = mt * 128 + tl;          // row of the super-tile this thread owns (TMEM lane tl of M-tile mt)
                if (ph == 0) {
                    // ---- epilogue 1: t1 = leaky(conv1 + b1), zero outside the utterance, as split-fp16 planes in smem ----
                    if (tile > 0) {                    // the previous tile's output tile aliases T1: its TMA store must have read it
                        if (tid < 2 * G) bulk_wait_read0();
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                    }
                    const int tr = it.t0 - p.pad2 + i;
                    const bool valid = tr >= 0 && tr < it.len;
                    const float isc = p.isc1;
                    uint8_t* dst = t1 + (size_t)i * 16;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float y = fmaf(racc[8 * g + j], isc, sbias[8 * g + j]);
                            y = y < 0.f ? y * 0.1f : y;
                            ovf |= valid && fabsf(y) > 8000.f;
                            v[j] = valid ? y : 0.f;
                        }
                        uint4 hi, lo;
                        split8(v, hi, lo);
                        *reinterpret_cast<uint4*>(dst + (size_t)g * XR2 * 16) = hi;
                        if (!p.mode) *reinterpret_cast<uint4*>(dst + (size_t)(G + g) * XR2 * 16) = lo;
                    }
                    fence_proxy_async();
                    mbar_arrive(t1_full);
                    RB_TS(2 + mt, tile * 8 + 2);
                } else {
                    // ---- epilogue 2: x' = act(conv2 + b2 + x), x from the x tile in smem; -> output tile -> TMA store ----
                    const int tr = it.t0 + i;
                    const bool valid = i < OV && tr < it.len;
                    const float isc = p.isc2;
                    const uint8_t* xs = abuf + (size_t)(buf * 2 + mt) * a_tile + (size_t)(tl + p.pad2 + p.pad1) * 16;
                    const float rinv = p.in_slope != 0.f ? 1.0f / p.in_slope : 1.0f;
                    mbar_wait(&a_full[buf], (uint32_t)((p.abufs == 2 ? (tile >> 1) : tile) & 1));   // complete long ago: acquire the TMA's writes
                    // The output tile [2G][OV][16 B] aliases T1, which conv2's MMAs of BOTH M-tiles read (M-tile 0's rows reach
                    // into the second half and the two layouts interleave): every epilogue thread has passed its own
                    // m_full / c_full wait here, so after this barrier all of conv2 has retired.
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    RB_TS(2 + mt, tile * 8 + 5);
                    uint8_t* dst = t1 + (size_t)i * 16;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const uint4 xh = *reinterpret_cast<const uint4*>(xs + (size_t)g * XR1 * 16);
                        const uint4 xl = *reinterpret_cast<const uint4*>(xs + (size_t)(G + g) * XR1 * 16);
                        const uint32_t hh[4] = {xh.x, xh.y, xh.z, xh.w}, ll[4] = {xl.x, xl.y, xl.z, xl.w};
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hh[j]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ll[j]));
                            float x0 = (fh.x + fl.x) * (1.0f / TC_ASCALE), x1 = (fh.y + fl.y) * (1.0f / TC_ASCALE);
                            x0 = x0 < 0.f ? x0 * rinv : x0;
                            x1 = x1 < 0.f ? x1 * rinv : x1;
                            float y0 = fmaf(racc[8 * g + 2 * j], isc, sbias[C + 8 * g + 2 * j]) + x0;
                            float y1 = fmaf(racc[8 * g + 2 * j + 1], isc, sbias[C + 8 * g + 2 * j + 1]) + x1;
                            if (p.out_act == ACT_LEAKY) {
                                y0 = y0 < 0.f ? y0 * p.out_slope : y0;
                                y1 = y1 < 0.f ? y1 * p.out_slope : y1;
                            }
                            ovf |= valid && (fabsf(y0) > 8000.f || fabsf(y1) > 8000.f);
                            o[2 * j] = valid ? y0 : 0.f;
                            o[2 * j + 1] = valid ? y1 : 0.f;
                        }
                        if (i < OV) {
                            uint4 hi, lo;
                            split8(o, hi, lo);
                            *reinterpret_cast<uint4*>(dst + (size_t)g * OV * 16) = hi;
                            *reinterpret_cast<uint4*>(dst + (size_t)(G + g) * OV * 16) = lo;
                        }
                    }
                    // the x tile has been read: after the barrier below it may be refilled (its other readers, conv1's MMAs,
                    // retired long ago)
                    fence_proxy_async();
                    RB_TS(2 + mt, tile * 8 + 6);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (tid < 2 * G) {        // one bulk store per (plane, 16-byte channel group): exactly the rows of this utterance
                        const int nrows = min(OV, it.len - it.t0);
                        __half* gdst = p.outp.base + ((size_t)tid * p.outp.rows_p + (size_t)(it.prow_u + it.t0)) * 8;
                        bulk_s2g(gdst, t1 + (size_t)tid * OV * 16, (uint32_t)nrows * 16);
                        bulk_commit();
                    }
                    if (tl == 0) mbar_arrive(&a_empty[buf]);
                    RB_TS(2 + mt, tile * 8 + 7);
                }
            }
        }
        if (tid < 2 * G) bulk_wait_all0();
        if (ovf && p.flags) atomicOr(p.flags, 1u);
    } else if (warp == 8 || warp == 9) {
        // ================= MMA issuers ================================================================
        if (p.trace && blockIdx.x == 0 && lane == 0) rtr = p.trace;
        const int role = warp - 8;                    // mode 0: 0 = main (A_hi x [W_hi|W_lo]), 1 = corr (A_lo x W_hi); mode 1: M-tile `role`
        const uint32_t id_main = (1u << 4) | ((uint32_t)((p.mode ? C : 2 * C) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id_corr = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t b_lbo = 2 * C * 16;            // bytes between the two 8-channel groups of a K-step (merged stage: 2C rows)
        const uint64_t b_bits = ((uint64_t)((b_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint32_t w_s = smem_u32(wst);
        uint32_t af_par[2] = {0, 0}, t1_par = 0, me_par[2] = {1, 1}, ce_par[2] = {1, 1};
        int bs = 0; uint32_t bph = 0;                 // ring slot / phase
        for (int tile = 0; rb_next(it, p.seg, OV); ++tile) {
            const int buf = p.abufs == 2 ? (tile & 1) : 0;
            for (int ph = 0; ph < 2; ++ph) {
                const int k = ph ? p.k2 : p.k1, dil = ph ? 1 : p.d1;
                const int XR = ph ? XR2 : XR1;
                RB_TS(role, tile * 8 + ph * 3);
                if (ph == 0) { mbar_wait_warp(&a_full[buf], af_par[buf]); af_par[buf] ^= 1; }
                else { mbar_wait_warp(t1_full, t1_par); t1_par ^= 1; }
                tc_fence_after();
                RB_TS(role, tile * 8 + ph * 3 + 1);
                const uint32_t a_lbo = (uint32_t)XR * 16;
                const uint64_t a_bits = ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
                const uint32_t a_lo = (uint32_t)G * a_lbo;            // lo plane follows the G groups of the hi plane
                const int NU = p.mode ? 1 : (k + UPT - 1) / UPT;
                const int sbase = p.resident ? (ph ? p.k1 : 0) : 0;
                for (int un = 0; un < NU; ++un) {
                    const int tap0 = un * UPT, tap1 = min(k, tap0 + UPT);
                    const int bs0 = bs; const uint32_t bph0 = bph;
                    for (int mt = 0; mt < 2; ++mt) {
                        if (p.mode && mt != role) continue;
                        const uint32_t a_s = ph ? smem_u32(t1) + (uint32_t)(mt * 128) * 16 : smem_u32(abuf) + (uint32_t)(buf * 2 + mt) * a_tile;
                        const uint32_t tmain = tmem + (uint32_t)(mt * 3 * C), tcorr = tmain + 2 * C;
                        if (p.mode || role == 0) { mbar_wait_warp(&m_empty[mt], me_par[mt]); me_par[mt] ^= 1; tc_fence_after(); }
                        else if (un == 0) { mbar_wait_warp(&c_empty[mt], ce_par[mt]); ce_par[mt] ^= 1; tc_fence_after(); }
                        if (!p.resident) { bs = bs0; bph = bph0; }     // both M-tile passes walk the same ring slots
                        for (int tap = tap0; tap < tap1; ++tap) {
                            int s;
                            if (p.resident) {
                                s = sbase + tap;
                                if (tile == 0 && (mt == 0 || p.mode)) { mbar_wait_warp(&b_full[s], 0); tc_fence_after(); }
                            } else {
                                s = bs;
                                if (mt == 0 || p.mode) { mbar_wait_warp(&b_full[s], bph); tc_fence_after(); }
                                if (++bs == p.nb) { bs = 0; bph ^= 1; }
                            }
                            const uint64_t da = a_bits | (uint64_t)(((a_s + (uint32_t)(tap * dil) * 16) & 0x3FFFFu) >> 4);
                            const uint64_t db = b_bits | (uint64_t)(((w_s + (uint32_t)s * stage) & 0x3FFFFu) >> 4);
                            if (elect_one()) {
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) {
                                    const uint64_t a = da + (uint32_t)(ks * ((2 * a_lbo) >> 4)), b = db + (uint32_t)(ks * ((2 * b_lbo) >> 4));
                                    if (p.mode || role == 0) tc_mma_f16(tmain, a, b, id_main, (tap == tap0 && ks == 0) ? 0u : 1u);
                                    else tc_mma_f16(tcorr, a + (a_lo >> 4), b, id_corr, (un == 0 && tap == 0 && ks == 0) ? 0u : 1u);
                                }
                                if (!p.resident && (mt == 1 || p.mode)) tc_commit(&b_empty[s]);
                            }
                            __syncwarp();
                        }
                        if (elect_one()) {
                            if (p.mode || role == 0) tc_commit(&m_full[mt]);
                            else if (un == NU - 1) tc_commit(&c_full[mt]);
                        }
                        __syncwarp();
                    }
                }
                RB_TS(role, tile * 8 + ph * 3 + 2);
            }
        }
    } else if (warp == 10) {
        // ================= x-tile producer (TMA) ======================================================
        if (lane == 0) {
            uint32_t ae_par[2] = {1, 1};
            for (int tile = 0; rb_next(it, p.seg, OV); ++tile) {
                const int buf = p.abufs == 2 ? (tile & 1) : 0;
                mbar_wait(&a_empty[buf], ae_par[buf]); ae_par[buf] ^= 1;
                const long long r0 = it.prow_u + it.t0 - p.pad2 - p.pad1;      // >= prow_u - TC_GAP >= 0
                mbar_expect_tx(&a_full[buf], 2 * a_tile);
                tma_load_3d(abuf + (size_t)(buf * 2 + 0) * a_tile, &imap, 0, (int)r0, 0, &a_full[buf]);
                tma_load_3d(abuf + (size_t)(buf * 2 + 1) * a_tile, &imap, 0, (int)r0 + 128, 0, &a_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ================= weight producer (bulk copies) ==============================================
        if (lane == 0) {
            if (p.resident) {
                RbTile probe = it;
                if (rb_next(probe, p.seg, OV)) {
                    for (int s = 0; s < p.k1 + p.k2; ++s) {
                        const uint8_t* src = s < p.k1 ? reinterpret_cast<const uint8_t*>(p.w1) + (size_t)s * stage
                                                      : reinterpret_cast<const uint8_t*>(p.w2) + (size_t)(s - p.k1) * stage;
                        mbar_expect_tx(&b_full[s], stage);
                        bulk_g2s(wst + (size_t)s * stage, src, stage, &b_full[s]);
                    }
                }
            } else {
                int s = 0; uint32_t ph = 1;
                while (rb_next(it, p.seg, OV)) {
                    for (int c = 0; c < 2; ++c) {
                        const int k = c ? p.k2 : p.k1;
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(c ? p.w2 : p.w1);
                        for (int tap = 0; tap < k; ++tap) {
                            mbar_wait(&b_empty[s], ph);
                            mbar_expect_tx(&b_full[s], stage);
                            bulk_g2s(wst + (size_t)s * stage, src + (size_t)tap * stage, stage, &b_full[s]);
                            if (++s == p.nb) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols));
    }
}

// ---------------------------------------------------------------------------------------------
// small companions
// ---------------------------------------------------------------------------------------------
// zero the TC_GAP rows before and after every utterance of a planes tensor (the fused kernel's TMA stores only ever write
// rows >= 0 of an utterance; rows past its end are written as zeros by the tile that straddles it)
__global__ void __launch_bounds__(256) planes_gap_zero_kernel(Planes pl, Seg seg) {
    const int u = blockIdx.y;
    const int len = seg_len(seg, u);
    const long long prow = planes_row(seg, u);
    const int groups = 2 * (pl.C / 8);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (group, side, row)
    if (idx >= groups * 2 * TC_GAP) return;
    const int gq = idx / (2 * TC_GAP), r = idx - gq * 2 * TC_GAP;
    const long long row = r < TC_GAP ? prow - TC_GAP + r : prow + len + (r - TC_GAP);
    *reinterpret_cast<uint4*>(pl.base + ((size_t)gq * pl.rows_p + (size_t)row) * 8) = make_uint4(0, 0, 0, 0);
}

// MRF mean (Generator_MS.cpp:177-196 / Generator_hifigan.cpp:154-173: xs = rb0(x); xs += rb1(x); xs += rb2(x); x = xs / n) over the
// planes the fused ResBlock1 branches produced (identity activation) -> fp32 rows [rows][C]; optionally also fp32 of one
// planes tensor alone (n == 1: test hook / single-kernel MRF)
__global__ void __launch_bounds__(256) mrf_combine_kernel(Planes a, Planes b, Planes c, int n, float div, Seg seg, float* __restrict__ y, int ldy,
                                                          float unleaky) {
    const int u = blockIdx.y;
    const int len = seg_len(seg, u);
    const int seg0 = seg_start(seg, u);
    const long long prow = planes_row(seg, u);
    const int G = a.C / 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // rows fastest inside a group: coalesced plane reads
    if (idx >= len * G) return;
    const int gq = idx / len, r = idx - gq * len;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    const Planes* ps[3] = {&a, &b, &c};
    for (int q = 0; q < n; ++q) {
        const Planes& pl = *ps[q];
        const size_t o = ((size_t)gq * pl.rows_p + (size_t)(prow + r)) * 8;
        const uint4 h = *reinterpret_cast<const uint4*>(pl.base + o);
        const uint4 l = *reinterpret_cast<const uint4*>(pl.base + (size_t)G * pl.rows_p * 8 + o);
        const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hh[j]));
            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ll[j]));
            float x0 = (fh.x + fl.x) * (1.0f / TC_ASCALE), x1 = (fh.y + fl.y) * (1.0f / TC_ASCALE);
            if (unleaky != 0.f) { x0 = x0 < 0.f ? x0 / unleaky : x0; x1 = x1 < 0.f ? x1 / unleaky : x1; }
            s[2 * j] = q == 0 ? x0 : s[2 * j] + x0;
            s[2 * j + 1] = q == 0 ? x1 : s[2 * j + 1] + x1;
        }
    }
    float* d = y + (size_t)(seg0 + r) * ldy + gq * 8;
    if (div != 1.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] /= div;
    }
    reinterpret_cast<float4*>(d)[0] = make_float4(s[0], s[1], s[2], s[3]);
    reinterpret_cast<float4*>(d)[1] = make_float4(s[4], s[5], s[6], s[7]);
}

// ---------------------------------------------------------------------------------------------
// host: packing, planning, launch
// ---------------------------------------------------------------------------------------------
inline void rb_prepare_weights(RbWeights& r, const float* w /*[k][C][CoutW]*/, int k, int C, int CoutW, int dil, int pad, const float* bias_dev,
                               std::vector<void*>& owned) {
    r.ok = false;
    if (C != 32 && C != 64) return;
    TcWeights t;
    std::vector<__half> buf;
    if (!tc_pack_weights_host(t, w, k, C, C, CoutW, buf, 8, /*force_merge=*/true)) return;
    if (!t.merge || t.NC != C || t.KC != C || t.kchunks != 1 || t.nchunks != 1) return;
    void* d = nullptr;
    if (cudaMalloc(&d, buf.size() * sizeof(__half)) != cudaSuccess) return;
    owned.push_back(d);
    if (cudaMemcpy(d, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return;
    r.packed = (__half*)d; r.inv_scale = t.inv_scale; r.k = k; r.dil = dil; r.pad = pad; r.bias = bias_dev;
    r.ok = true;
}
inline bool rb_pair_eligible(const RbWeights& a, const RbWeights& b) {
    if (!a.ok || !b.ok) return false;
    if (b.dil != 1 || 2 * b.pad != b.k - 1 || 2 * a.pad != (a.k - 1) * a.dil) return false;
    if (a.pad + b.pad > TC_GAP || 128 + (a.k - 1) * a.dil > 256 || b.pad > a.pad) return false;
    if (a.k + b.k > RB_MAX_STAGES || b.k > 33) return false;
    return true;
}
struct RbPlan { size_t smem; int abufs, resident, nb; };
inline RbPlan rb_plan(int C, const RbWeights& a, const RbWeights& b, size_t budget = 225 * 1024) {
    RbPlan pl;
    const size_t xr1 = 128 + (size_t)(a.k - 1) * a.dil, xr2 = 256 + b.k - 1;
    const size_t a_tile = (size_t)C * xr1 * 4, t1 = (size_t)C * xr2 * 4, stage = (size_t)4 * C * C;
    const size_t misc = (13 + 2 * RB_MAX_STAGES) * 8 + 16 + 2 * C * 4 + 256;
    const size_t wres = (size_t)(a.k + b.k) * stage;
    static const int e_ab = getenv("STTS_RB_ABUFS") ? atoi(getenv("STTS_RB_ABUFS")) : 0;
    static const int e_res = getenv("STTS_RB_RES") ? atoi(getenv("STTS_RB_RES")) : -1;
    static const int e_nb = getenv("STTS_RB_NB") ? atoi(getenv("STTS_RB_NB")) : 0;
    const int upt = std::max(1, 8 / (C / 16));
    const size_t ring_min = (size_t)(upt + 1) * stage;
    // preference: resident weights + double-buffered x tiles > resident + single > ring (>= 4 stages) + double > ring + single
    pl.abufs = 1; pl.resident = 0; pl.nb = upt + 1;
    if (2 * 2 * a_tile + t1 + wres + misc <= budget) { pl.abufs = 2; pl.resident = 1; }
    else if (2 * a_tile + t1 + wres + misc <= budget) { pl.abufs = 1; pl.resident = 1; }
    else if (2 * 2 * a_tile + t1 + ring_min + misc <= budget) pl.abufs = 2;
    if (e_ab == 1 || e_ab == 2) pl.abufs = e_ab;
    if (e_res == 0) pl.resident = 0;
    if (pl.resident) pl.nb = a.k + b.k;
    else {
        const size_t base = (size_t)pl.abufs * 2 * a_tile + t1 + misc;
        size_t room = budget > base ? (budget - base) / stage : 0;
        pl.nb = (int)std::min<size_t>(std::min<size_t>(room, 8), RB_MAX_STAGES);
        if (e_nb > 0) pl.nb = std::min(e_nb, RB_MAX_STAGES);
        if (pl.nb < upt + 1 && pl.abufs == 2) {      // make room by dropping the second x-tile buffer
            pl.abufs = 1;
            const size_t base1 = 2 * a_tile + t1 + misc;
            room = budget > base1 ? (budget - base1) / stage : 0;
            pl.nb = (int)std::min<size_t>(std::min<size_t>(room, 8), RB_MAX_STAGES);
        }
    }
    pl.smem = (size_t)pl.abufs * 2 * a_tile + t1 + (size_t)pl.nb * stage + misc;
    return pl;
}
inline bool rb_make_map(CUtensorMap* m, const Planes& pl, int box_rows) {
    cuuint64_t dims[3] = {8, (cuuint64_t)pl.rows_p, (cuuint64_t)(2 * (pl.C / 8))};
    cuuint64_t strides[2] = {16, (cuuint64_t)pl.rows_p * 16};
    cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)(2 * (pl.C / 8))};
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
// per-device one-time setup (dynamic shared memory limit): call from stts_engine::build() on the engine's device
inline cudaError_t rb_device_setup() {
    cudaError_t e = cudaFuncSetAttribute(rb_pair_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(rb_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
// x planes (act applied by the producer) -> x' planes.  Returns 1 (launches), < 0 on error.
inline int rb_pair_launch(int C, const RbWeights& a, const RbWeights& b, const Planes& in, const Planes& out, Seg seg, int nseg, int maxlen,
                          float in_slope, int out_act, float out_slope, int mode, int sms, unsigned int* flags, cudaStream_t stream) {
    if (!rb_pair_eligible(a, b) || in.C != C || out.C != C) return -3;
    RbP p;
    p.seg = seg;
    p.w1 = a.packed; p.w2 = b.packed; p.b1 = a.bias; p.b2 = b.bias; p.isc1 = a.inv_scale; p.isc2 = b.inv_scale;
    p.k1 = a.k; p.d1 = a.dil; p.pad1 = a.pad; p.k2 = b.k; p.pad2 = b.pad;
    p.xr1 = 128 + (a.k - 1) * a.dil; p.xr2 = 256 + b.k - 1; p.ov = 256 - (b.k - 1);
    static const int env_us = getenv("STTS_TC_USTEPS") ? atoi(getenv("STTS_TC_USTEPS")) : 0;
    p.usteps = env_us > 0 ? env_us : 8;
    p.mode = mode;
    const RbPlan pl = rb_plan(C, a, b);
    const int upt_l = std::max(1, p.usteps / (C / 16));
    if (pl.smem > 227 * 1024 || (!pl.resident && pl.nb < std::min(upt_l, std::max(a.k, b.k)) + 1)) return -4;
    p.abufs = pl.abufs; p.resident = pl.resident; p.nb = pl.nb;
    p.in_slope = in_slope; p.out_act = out_act; p.out_slope = out_slope;
    p.tmem_cols = C == 32 ? 256 : 512;
    p.flags = flags;
    const int gx = (maxlen + p.ov - 1) / p.ov;
    const long long W = (long long)gx * nseg;
    if (W <= 0 || W > 0x7fffffffLL) return -1;
    p.gx = gx; p.work_items = (int)W;
    alignas(64) CUtensorMap imap;
    if (!rb_make_map(&imap, in, p.xr1)) return -1;
    p.outp = out;
    static const int env_grid = getenv("STTS_RB_GRID") ? atoi(getenv("STTS_RB_GRID")) : 0;
    int ctas = env_grid > 0 ? env_grid : sms;
    if ((long long)ctas > W) ctas = (int)W;
    static const int env_verbose = getenv("STTS_TC_VERBOSE") ? atoi(getenv("STTS_TC_VERBOSE")) : 0;
    if (env_verbose > 0) {
        static int left = env_verbose;
        if (left > 0) { --left; fprintf(stderr, "rb_pair: C=%d k=%d d=%d mode=%d items=%d ctas=%d smem=%zu abufs=%d resident=%d nb=%d ov=%d\n", C, a.k, a.dil, mode, p.work_items, ctas, pl.smem, pl.abufs, pl.resident, pl.nb, p.ov); }
    }
    static long long* trace_buf = nullptr;
    static const int env_trace = getenv("STTS_RB_TRACE") ? atoi(getenv("STTS_RB_TRACE")) : 0;   // k to trace (first matching launches per C)
    static int trace_left[2] = {1, 1};
    const bool do_trace = env_trace && a.k == env_trace && a.dil == (a.k == 3 ? 1 : (a.k == 7 ? 3 : 5)) && trace_left[C == 64] > 0 && maxlen > 1024;
    if (do_trace && !trace_buf) cudaMalloc(&trace_buf, 6 * 1024 * 8);
    if (do_trace) cudaMemsetAsync(trace_buf, 0, 6 * 1024 * 8, stream);
    p.trace = do_trace ? trace_buf : nullptr;
    if (C == 32) rb_pair_kernel<32><<<ctas, RB_THREADS, pl.smem, stream>>>(p, imap);
    else rb_pair_kernel<64><<<ctas, RB_THREADS, pl.smem, stream>>>(p, imap);
    if (do_trace) {   // dump the traced CTA's timeline (debug tool; synchronises)
        --trace_left[C == 64];
        std::vector<long long> h(6 * 1024);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), trace_buf, h.size() * 8, cudaMemcpyDeviceToHost);
        long long t0 = 0;
        for (auto v : h) if (v && (!t0 || v < t0)) t0 = v;
        static const char* names[6] = {"iss0", "iss1", "set0", "set1", "xprod", "wprod"};
        fprintf(stderr, "RBTRACE C=%d k=%d d=%d mode=%d items=%d ctas=%d abufs=%d resident=%d nb=%d\n", C, a.k, a.dil, mode, p.work_items, ctas, pl.abufs, pl.resident, pl.nb);
        for (int r = 0; r < 4; ++r) {
            fprintf(stderr, " %s:", names[r]);
            for (int i = 0; i < 64; ++i) if (h[r * 1024 + i]) fprintf(stderr, " %d:%lld", i, h[r * 1024 + i] - t0);
            fprintf(stderr, "\n");
        }
    }
    return 1;
}

}  // namespace stts
