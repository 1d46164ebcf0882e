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

constexpr int RB_THREADS = 704;     // warps 0-15: four epilogue sets (M-tile x column half); 16-19: MMA issuers (kind x M-tile); 20: weight producer; 21: x-tile producer
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
    const int2* tiles;               // [work_items] (utterance, first row) of every super-tile, built by rb_tiles_kernel
    int work_items;                  // number of super-tiles: a kernel PARAMETER, so every role's tile loop has uniform bounds
                                     // (the MMA issuers keep their descriptors in uniform registers only in provably uniform loops)
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
    Planes inp; int bulk_in;         // input planes; 1: x tiles by 1-D bulk copies (planes_tile_g2s), 0: tensor-map boxes (STTS_TILE_TMA=1)
    Planes outp;                     // destination planes (written with 1-D bulk stores of the exact valid rows)
    unsigned int* flags;             // bit 0: an activation exceeded the fp16 range of the split (|8 x| > 65504)
    int dbg;                         // experiments (STTS_RB_DBG): bit 0 = epilogue / producer mbarrier polls back off with nanosleep
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
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// lane 0 polls, the warp re-converges, then every lane observes the (already complete) phase itself
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* b, uint32_t parity, unsigned ns) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!done) __nanosleep(ns);
    }
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_all(uint64_t* b, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(b, parity);
    __syncwarp();
    mbar_wait(b, parity);
}

struct RbTile { int u, seg0, len, t0; long long prow_u; };
__device__ __forceinline__ RbTile rb_tile_at(const Seg& seg, const int2* tiles, int w) {
    RbTile it;
    const int2 e = __ldg(tiles + w);
    it.u = e.x; it.t0 = e.y;
    it.len = seg_len(seg, e.x);
    it.seg0 = seg_start(seg, e.x);
    it.prow_u = planes_row(seg, e.x);
    return it;
}
__device__ __forceinline__ RbTile rb_tile(const RbP& p, int w) { return rb_tile_at(p.seg, p.tiles, w); }
// (utterance, first row) of every super-tile of `ov` rows, utterance-major; one block
__global__ void __launch_bounds__(256) rb_tiles_kernel(Seg seg, int nseg, int ov, int2* __restrict__ out, int cap) {
    __shared__ int scan[256];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int u0 = 0; u0 < nseg; u0 += 256) {
        const int u = u0 + threadIdx.x;
        const int n = u < nseg ? (seg_len(seg, u) + ov - 1) / ov : 0;
        scan[threadIdx.x] = n;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int v = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
            __syncthreads();
            scan[threadIdx.x] += v;
            __syncthreads();
        }
        const int first = base + scan[threadIdx.x] - n;
        for (int x = 0; x < n; ++x)
            if (first + x < cap) out[first + x] = make_int2(u, x * ov);
        __syncthreads();
        if (threadIdx.x == 255) base += scan[255];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// MMA issuer (one warp per ROLE).  Everything that feeds a tcgen05.mma operand is derived from kernel parameters,
// blockIdx and loop counters with parameter bounds only, and ROLE is a template constant: ptxas then keeps the descriptors
// in uniform registers (a data-dependent tile loop cost ~14 R2UR moves and 250-300 cycles per MMA here).
//   mode 0: ROLE 0 issues A_hi x [W_hi | W_lo] (N = 2C) into main[mt], ROLE 1 issues A_lo x W_hi (N = C) into corr[mt];
//   mode 1: ROLE r issues A_hi x W_hi (N = C) for M-tile r.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int rb_nmain(int C) { return C <= 32 ? 2 : 1; }
template <int C, int KIND>
__device__ __forceinline__ void rb_issuer(const RbP& p, const int mt, const uint32_t abuf_s, const uint32_t t1_s, const uint32_t w_s, const uint32_t tmem,
                                          uint64_t* a_full, uint64_t* m_full, uint64_t* m_empty, uint64_t* c_full, uint64_t* c_empty,
                                          uint64_t* t1_full, uint64_t* b_full, uint64_t* b_empty, long long* rtr) {
    // ONE thread per (KIND, M-tile) runs a whole issue loop: four issuing threads per CTA.  A single thread sustains one
    // MMA per ~100 cycles in this loop (R2UR moves + the per-instruction election wrapper ptxas emits in divergent code),
    // the tensor pipe retires a small-N MMA in 40-65 cycles, so four streams keep it busy and hide each other's waits.
    //   mode 0: KIND 0 = A_hi x [W_hi | W_lo] (N = 2C) into main[mt] in promotion units, KIND 1 = A_lo x W_hi (N = C) into corr[mt];
    //   mode 1: KIND 0 = even taps, KIND 1 = odd taps of A_hi x W_hi (N = C), into main[mt] / corr[mt] (summed by the epilogue).
    constexpr int G = C / 8, KS = C / 16;
    const int mode = p.mode;
    const uint32_t a_tile = (uint32_t)C * p.xr1 * 4;
    const uint32_t stage16 = (uint32_t)(4 * C * C) >> 4;                 // weight stage in 16-byte units
    const int UPT = mode ? 64 : max(1, p.usteps / KS);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(((mode || KIND == 1) ? C : 2 * C) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t b_lbo = 2 * C * 16;        // bytes between the two 8-channel groups of a K-step (merged stage: 2C rows)
    constexpr uint32_t b_k16 = (2 * b_lbo) >> 4;
    const uint64_t b_desc0 = ((uint64_t)((b_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46) | (uint64_t)((w_s & 0x3FFFFu) >> 4);
    uint32_t af_par0 = 0, af_par1 = 0, t1_par = 0, e_par0 = 1, e_par1 = 1;
    int bs = 0; uint32_t bph = 0;                 // ring slot / phase
    const int W = p.work_items, step = gridDim.x;
    // TMEM per M-tile: main[NMAIN] (2C columns each) then corr (C columns).  With two main accumulators (C = 32: TMEM has the
    // room) the KIND 0 issuer runs two promotion units ahead of the drains, so conv1 of the next tile overlaps epilogue 2.
    constexpr int NMAIN = rb_nmain(C);
    constexpr uint32_t MT_COLS = (2 * NMAIN + 1) * C;
    const uint32_t d_base = tmem + (uint32_t)mt * MT_COLS;
    uint32_t ucount = 0;                            // units issued by this thread (selects the main accumulator)
    const int tstep = mode ? 2 : 1, toff = mode ? KIND : 0;     // mode 1: this issuer's taps are toff, toff + 2, ...
    int tile = 0;
#ifdef STTS_TC_TRACE_BUILD
    int fine = 0;
#endif
    for (int w = blockIdx.x; w < W; w += step, ++tile) {
        const int buf = p.abufs == 2 ? (tile & 1) : 0;
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
            const int k = ph ? p.k2 : p.k1;
            const uint32_t dil = ph ? 1u : (uint32_t)p.d1;
            const int XR = ph ? p.xr2 : p.xr1;
            RB_TS(KIND, tile * 8 + ph * 3);
            if (ph == 0) {
                if (buf) { mbar_wait(&a_full[1], af_par1); af_par1 ^= 1; } else { mbar_wait(&a_full[0], af_par0); af_par0 ^= 1; }
            } else { mbar_wait(t1_full, t1_par); t1_par ^= 1; }
            tc_fence_after();
            RB_TS(KIND, tile * 8 + ph * 3 + 1);
            const uint32_t a_lbo = (uint32_t)XR * 16;
            const uint64_t a_bits = ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint32_t a_k16 = (2 * a_lbo) >> 4;
            const uint32_t a_plane = (mode == 0 && KIND == 1) ? (uint32_t)G * a_lbo : 0u;      // mode 0 corr reads the lo plane
            const uint32_t a_s = (ph ? t1_s + (uint32_t)(mt * 128) * 16 : abuf_s + (uint32_t)(buf * 2 + mt) * a_tile) + a_plane;
            const int NU = (mode || KIND == 1) ? 1 : (k + UPT - 1) / UPT;       // promotion units (main accumulator of mode 0 only)
            const int sbase = (p.resident && ph) ? p.k1 : 0;
#pragma unroll 1
            for (int un = 0; un < NU; ++un) {
                const int tap0 = (NU == 1 ? 0 : un * UPT) + toff, tap1 = NU == 1 ? k : min(k, un * UPT + UPT);
                const uint32_t mb = (KIND == 0 && NMAIN == 2) ? (ucount & 1u) : 0u;
                ++ucount;
                uint64_t* full_bar = KIND == 0 ? &m_full[mt * 2 + mb] : &c_full[mt];
                uint64_t* empty_bar = KIND == 0 ? &m_empty[mt * 2 + mb] : &c_empty[mt];
                const uint32_t d_t = d_base + (KIND == 1 ? (uint32_t)(2 * NMAIN * C) : mb * 2 * C);
                if (mb) { mbar_wait(empty_bar, e_par1); e_par1 ^= 1; } else { mbar_wait(empty_bar, e_par0); e_par0 ^= 1; }   // accumulator drained
                tc_fence_after();
                uint64_t da = a_bits | (uint64_t)(((a_s + (uint32_t)tap0 * dil * 16) & 0x3FFFFu) >> 4);
                uint32_t acc = 0u;
#ifdef STTS_TC_TRACE_BUILD
                if (tile == 1 && mt == 0 && fine < 250) { RB_TS(4 + KIND, fine * 2); }
#endif
                if (p.resident) {
                    if (tile == 0) {
                        for (int tap = tap0; tap < tap1; tap += tstep) mbar_wait(&b_full[sbase + tap], 0);
                        tc_fence_after();
                    }
                    uint64_t db = b_desc0 + (uint32_t)(sbase + tap0) * stage16;
#pragma unroll 1
                    for (int tap = tap0; tap < tap1; tap += tstep) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) tc_mma_f16(d_t, da + (uint32_t)(ks * a_k16), db + (uint32_t)(ks * b_k16), idesc, ks == 0 ? acc : 1u);
                        acc = 1u;
                        da += dil * (uint32_t)tstep; db += stage16 * (uint32_t)tstep;
                    }
                } else {
                    // ring: every issuer walks all the slots of the unit, waits / issues / frees only at its own taps
                    const int r0 = NU == 1 ? 0 : un * UPT;
#pragma unroll 1
                    for (int tap = r0; tap < tap1; ++tap) {
                        const int s = bs;
                        const uint32_t sph = bph;
                        if (++bs == p.nb) { bs = 0; bph ^= 1; }
                        if (mode && ((tap - toff) & 1)) continue;
                        mbar_wait(&b_full[s], sph);
                        tc_fence_after();
                        const uint64_t db = b_desc0 + (uint32_t)s * stage16;
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) tc_mma_f16(d_t, da + (uint32_t)(ks * a_k16), db + (uint32_t)(ks * b_k16), idesc, ks == 0 ? acc : 1u);
                        acc = 1u;
                        da += dil * (uint32_t)tstep;
                        tc_commit(&b_empty[s]);
                    }
                }
                tc_commit(full_bar);
#ifdef STTS_TC_TRACE_BUILD
                if (tile == 1 && mt == 0 && fine < 250) { RB_TS(4 + KIND, fine * 2 + 1); ++fine; }
#endif
            }
            RB_TS(KIND, tile * 8 + ph * 3 + 2);
        }
    }
}

// 16-column / 32-column TMEM loads without the wait (several in flight, one tcgen05.wait::ld)
template <int CH>
__device__ __forceinline__ void tc_ld_nowait(uint32_t taddr, uint32_t* r) {
    if constexpr (CH == 32) tc_ld32_nowait(taddr, r);
    else
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
}
// 8 values already in the x8 domain -> hi | lo fp16 planes (two packed conversions per pair; saturating)
__device__ __forceinline__ uint32_t cvt_h2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void split8_scaled(const float* v, uint4& hi, uint4& lo) {
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hh[i] = cvt_h2_sat(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hh[i]));
        ll[i] = cvt_h2_sat(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    }
    hi = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    lo = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

template <int C>
__global__ void __launch_bounds__(RB_THREADS, 1) rb_pair_kernel(const RbP p, const __grid_constant__ CUtensorMap imap) {
    constexpr int G = C / 8;          // 16-byte channel groups per plane
    constexpr int KS = C / 16;        // K = 16 MMA steps per tap
    constexpr int HS = C >= 32 ? 2 : 1;   // column halves: a row is shared by HS threads of different warps (C = 16: one thread per row,
                                          // epilogue warps 8-15 idle)
    constexpr int CH = C / HS;        // columns per epilogue thread
    constexpr int GH = G / HS;        // channel groups per epilogue thread
    extern __shared__ __align__(128) uint8_t rsm[];
    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a warp reduction (REDUX writes a uniform register: the role branches below are warp-uniform)
    const int warp = __reduce_max_sync(0xffffffffu, tid >> 5);
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
    uint64_t* m_full = bars + 4;         // [2 M-tiles][2 main accumulators]
    uint64_t* m_empty = bars + 8;        // [2][2]
    uint64_t* c_full = bars + 12;        // [2]
    uint64_t* c_empty = bars + 14;       // [2]
    uint64_t* t1_full = bars + 16;       // [1]
    uint64_t* b_full = bars + 17;        // [RB_MAX_STAGES]
    uint64_t* b_empty = b_full + RB_MAX_STAGES;   // [RB_MAX_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + RB_MAX_STAGES);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);   // [2][C], x8 domain

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2 * HS);      // one arrival per epilogue set (M-tile, column half)
            mbar_init(&m_full[2 * i], 1); mbar_init(&m_empty[2 * i], 4 * HS);      // one arrival per warp of the M-tile's sets
            mbar_init(&m_full[2 * i + 1], 1); mbar_init(&m_empty[2 * i + 1], 4 * HS);
            mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 4 * HS);
        }
        mbar_init(t1_full, 8 * HS);                                    // one arrival per active epilogue warp
        for (int s = 0; s < RB_MAX_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], p.mode ? 2 : 4); }   // issuers that read a stage
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * C) sbias[tid] = TC_ASCALE * (tid < C ? (p.b1 ? __ldg(p.b1 + tid) : 0.f) : (p.b2 ? __ldg(p.b2 + tid - C) : 0.f));
    for (uint32_t i = tid; i < t1_bytes / 16; i += RB_THREADS) reinterpret_cast<uint4*>(t1)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;      // per M-tile: main[NMAIN] (2C columns each: hi*hi | hi*lo) then corr (C columns)

    const int W = p.work_items, wstep = gridDim.x;
    long long* rtr = nullptr;
    const int UPT = p.mode ? 64 : max(1, p.usteps / KS);     // taps per promotion unit

    if (warp < 16 && (warp >> 3) >= HS) {
        // (C = 16: the second column-half warps have nothing to do)
    } else if (warp < 16) {
        // ================= promotion + epilogues: set (M-tile mt, column half hf), one row x CH columns per thread ==========
        const int wq = warp & 3, mt = (warp >> 2) & 1, hf = warp >> 3;
        const int tl = wq * 32 + lane;                 // TMEM lane = row inside the M-tile
        const int c0 = hf * CH, g0 = hf * GH;
        constexpr int NMAIN = rb_nmain(C);
        constexpr uint32_t MT_COLS = (2 * NMAIN + 1) * C;
        const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)mt * MT_COLS + (uint32_t)c0;
        const uint32_t tcorr = tbase + 2 * NMAIN * C;
        uint32_t ucount = 0, mf_par1 = 0;
        if (p.trace && blockIdx.x == 0 && tl == 0 && hf == 0) rtr = p.trace;
        uint32_t mf_par = 0, cf_par = 0;
        float amax = 0.f;                               // largest |8 x| converted (overflow check of the saturating split)
        float racc[CH];
        const float isc1 = p.isc1 * TC_ASCALE, isc2 = p.isc2 * TC_ASCALE;
        const float rinv = p.in_slope != 0.f ? 1.0f / p.in_slope : 1.0f;              // unleaky(s) = min(s, s * rinv)
        const float oslope = p.out_act == ACT_LEAKY ? p.out_slope : 1.0f;              // leaky(y) = max(y, y * slope)
        int tile = 0;
        for (int w = blockIdx.x; w < W; w += wstep, ++tile) {
            const RbTile it = rb_tile(p, w);
            const int buf = p.abufs == 2 ? (tile & 1) : 0;
            for (int ph = 0; ph < 2; ++ph) {
                const int k = ph ? p.k2 : p.k1;
                const int NU = p.mode ? 1 : (k + UPT - 1) / UPT;
                for (int un = 0; un < NU; ++un) {
                    const uint32_t mb = NMAIN == 2 ? (ucount & 1u) : 0u;
                    ++ucount;
                    const uint32_t tmain = tbase + mb * 2 * C;
                    if (mb) { mbar_wait_all(&m_full[mt * 2 + 1], mf_par1); mf_par1 ^= 1; } else { mbar_wait_all(&m_full[mt * 2], mf_par); mf_par ^= 1; }
                    tc_fence_after();
                    if (un == 0) RB_TS(2 + mt, tile * 8 + ph * 3);
#pragma unroll
                    for (int cb = 0; cb < CH; cb += 16) {      // two tcgen05.ld in flight per 16-column chunk, one wait
                        uint32_t v[16], x2[16];
                        tc_ld_nowait<16>(tmain + cb, v);
                        if (!p.mode) tc_ld_nowait<16>(tmain + C + cb, x2);
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
                    if (lane == 0) mbar_arrive(&m_empty[mt * 2 + mb]);
                }
                {
                    mbar_wait_all(&c_full[mt], cf_par); cf_par ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int cb = 0; cb < CH; cb += 16) {
                        uint32_t v[16];
                        tc_ld_nowait<16>(tcorr + cb, v);
                        tc_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) racc[cb + j] += __uint_as_float(v[j]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&c_empty[mt]);
                }
                RB_TS(2 + mt, tile * 8 + ph * 3 + 1);
                const int i = mt * 128 + tl;          // row of the super-tile this thread works on
                if (ph == 0) {
                    // ---- epilogue 1: t1 = leaky(conv1 + b1), zero outside the utterance, as split-fp16 planes in smem ----
                    if (tile > 0) {                    // the previous tile's output tile aliases T1: its bulk stores must have read it
                        if (tid < 2 * G) bulk_wait_read0();
                        asm volatile("bar.sync 1, %0;" ::"n"(256 * HS) : "memory");
                    }
                    const int tr = it.t0 - p.pad2 + i;
                    const bool valid = tr >= 0 && tr < it.len;
                    uint8_t* dst = t1 + (size_t)i * 16;
#pragma unroll
                    for (int gg = 0; gg < GH; ++gg) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float y = fmaf(racc[8 * gg + j], isc1, sbias[c0 + 8 * gg + j]);
                            y = fmaxf(y, y * 0.1f);
                            amax = fmaxf(amax, valid ? fabsf(y) : 0.f);
                            v[j] = y;
                        }
                        uint4 hi, lo;
                        split8_scaled(v, hi, lo);
                        if (!valid) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
                        *reinterpret_cast<uint4*>(dst + (size_t)(g0 + gg) * XR2 * 16) = hi;
                        if (!p.mode) *reinterpret_cast<uint4*>(dst + (size_t)(G + g0 + gg) * XR2 * 16) = lo;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t1_full);
                    RB_TS(2 + mt, tile * 8 + 2);
                } else {
                    // ---- epilogue 2: x' = act(conv2 + b2 + x), x from the x tile in smem; -> output tile -> bulk stores ----
                    const int tr = it.t0 + i;
                    const bool valid = i < OV && tr < it.len;
                    const uint8_t* xs = abuf + (size_t)(buf * 2 + mt) * a_tile + (size_t)(tl + p.pad2 + p.pad1) * 16;
                    mbar_wait(&a_full[buf], (uint32_t)((p.abufs == 2 ? (tile >> 1) : tile) & 1));   // complete long ago: acquire the TMA's writes
                    // The output tile [2G][OV][16 B] aliases T1, which conv2's MMAs of BOTH M-tiles read (M-tile 0's rows reach
                    // into the second half and the two layouts interleave): every epilogue thread has passed its own
                    // m_full / c_full wait here, so after this barrier all of conv2 has retired.
                    asm volatile("bar.sync 2, %0;" ::"n"(256 * HS) : "memory");
                    RB_TS(2 + mt, tile * 8 + 5);
                    uint8_t* dst = t1 + (size_t)i * 16;
#pragma unroll
                    for (int gg = 0; gg < GH; ++gg) {
                        const uint4 xh = *reinterpret_cast<const uint4*>(xs + (size_t)(g0 + gg) * XR1 * 16);
                        const uint4 xl = *reinterpret_cast<const uint4*>(xs + (size_t)(G + g0 + gg) * XR1 * 16);
                        const uint32_t hh[4] = {xh.x, xh.y, xh.z, xh.w}, ll[4] = {xl.x, xl.y, xl.z, xl.w};
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hh[j]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ll[j]));
                            float x0 = fh.x + fl.x, x1 = fh.y + fl.y;               // 8 * leaky(x)
                            x0 = fminf(x0, x0 * rinv);
                            x1 = fminf(x1, x1 * rinv);
                            float y0 = fmaf(racc[8 * gg + 2 * j], isc2, sbias[C + c0 + 8 * gg + 2 * j]) + x0;
                            float y1 = fmaf(racc[8 * gg + 2 * j + 1], isc2, sbias[C + c0 + 8 * gg + 2 * j + 1]) + x1;
                            y0 = fmaxf(y0, y0 * oslope);
                            y1 = fmaxf(y1, y1 * oslope);
                            amax = fmaxf(amax, valid ? fmaxf(fabsf(y0), fabsf(y1)) : 0.f);
                            o[2 * j] = y0; o[2 * j + 1] = y1;
                        }
                        if (i < OV) {
                            uint4 hi, lo;
                            split8_scaled(o, hi, lo);
                            if (!valid) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
                            *reinterpret_cast<uint4*>(dst + (size_t)(g0 + gg) * OV * 16) = hi;
                            *reinterpret_cast<uint4*>(dst + (size_t)(G + g0 + gg) * OV * 16) = lo;
                        }
                    }
                    // the x tile has been read: after the barrier below it may be refilled (its other readers, conv1's MMAs,
                    // retired long ago)
                    fence_proxy_async();
                    RB_TS(2 + mt, tile * 8 + 6);
                    asm volatile("bar.sync 1, %0;" ::"n"(256 * HS) : "memory");
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
        if (amax > 65000.f && p.flags) atomicOr(p.flags, 1u);
    } else if (warp < 20) {
        // ================= MMA issuers: warp 16 + 2 KIND + mt, one thread each ========================
        if (lane == 0) {
            const int kind = (warp - 16) >> 1, imt = (warp - 16) & 1;
            if (p.trace && blockIdx.x == 0 && imt == 0) rtr = p.trace;
            if (kind == 0) rb_issuer<C, 0>(p, imt, smem_u32(abuf), smem_u32(t1), smem_u32(wst), tmem, a_full, m_full, m_empty, c_full, c_empty, t1_full, b_full, b_empty, rtr);
            else rb_issuer<C, 1>(p, imt, smem_u32(abuf), smem_u32(t1), smem_u32(wst), tmem, a_full, m_full, m_empty, c_full, c_empty, t1_full, b_full, b_empty, rtr);
        }
        __syncwarp();
    } else if (warp == 20) {
        // ================= weight producer: one thread, blocking loop (resident: every stage once; else the ring) ==========
        // (One thread polling both streams cooperatively costs ~100 dependent instructions per stage — pc_fused.cuh measured
        //  1.2k cycles per stage for that pattern; two simple blocking loops in two warps cost a few dozen.)
        if (lane == 0) {
            if (p.resident) {
                if ((int)blockIdx.x < W)
                    for (int s2 = 0; s2 < p.k1 + p.k2; ++s2) {
                        const uint8_t* src = s2 < p.k1 ? reinterpret_cast<const uint8_t*>(p.w1) + (size_t)s2 * stage
                                                       : reinterpret_cast<const uint8_t*>(p.w2) + (size_t)(s2 - p.k1) * stage;
                        mbar_expect_tx(&b_full[s2], stage);
                        bulk_g2s(wst + (size_t)s2 * stage, src, stage, &b_full[s2]);
                    }
            } else {
                int ws_ = 0; uint32_t wph = 1;
#pragma unroll 1
                for (int ww = blockIdx.x; ww < W; ww += wstep)
#pragma unroll 1
                    for (int wc = 0; wc < 2; ++wc) {
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(wc ? p.w2 : p.w1);
                        const int kk = wc ? p.k2 : p.k1;
#pragma unroll 1
                        for (int wtap = 0; wtap < kk; ++wtap) {
                            mbar_wait(&b_empty[ws_], wph);
                            mbar_expect_tx(&b_full[ws_], stage);
                            bulk_g2s(wst + (size_t)ws_ * stage, src + (size_t)wtap * stage, stage, &b_full[ws_]);
                            if (++ws_ == p.nb) { ws_ = 0; wph ^= 1; }
                        }
                    }
            }
        }
        __syncwarp();
    } else {
        // ================= x-tile producer (warp 21): one thread, blocking =====================================================
        if (lane == 0) {
            uint32_t ae_par0 = 1, ae_par1 = 1;
            int xtile = 0;
#pragma unroll 1
            for (int xw = blockIdx.x; xw < W; xw += wstep, ++xtile) {
                const int buf = p.abufs == 2 ? (xtile & 1) : 0;
                if (buf) { mbar_wait(&a_empty[1], ae_par1); ae_par1 ^= 1; } else { mbar_wait(&a_empty[0], ae_par0); ae_par0 ^= 1; }
                const RbTile it = rb_tile(p, xw);
                const long long r0 = it.prow_u + it.t0 - p.pad2 - p.pad1;      // >= prow_u - TC_GAP >= 0
                mbar_expect_tx(&a_full[buf], 2 * a_tile);
                if (p.bulk_in) {
                    planes_tile_g2s(abuf + (size_t)(buf * 2 + 0) * a_tile, p.inp, 0, 2 * (C / 8), r0, p.xr1, &a_full[buf]);
                    planes_tile_g2s(abuf + (size_t)(buf * 2 + 1) * a_tile, p.inp, 0, 2 * (C / 8), r0 + 128, p.xr1, &a_full[buf]);
                } else {
                    tma_load_3d(abuf + (size_t)(buf * 2 + 0) * a_tile, &imap, 0, (int)r0, 0, &a_full[buf]);
                    tma_load_3d(abuf + (size_t)(buf * 2 + 1) * a_tile, &imap, 0, (int)r0 + 128, 0, &a_full[buf]);
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 16) {
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
    if (C != 16 && C != 32 && C != 64) return;
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
    if (a.k + b.k > RB_MAX_STAGES || b.k > 33 || a.k < 2 || b.k < 2) return false;
    return true;
}
struct RbPlan { size_t smem; int abufs, resident, nb; };
inline RbPlan rb_plan(int C, const RbWeights& a, const RbWeights& b, size_t budget = 225 * 1024) {
    RbPlan pl;
    const size_t xr1 = 128 + (size_t)(a.k - 1) * a.dil, xr2 = 256 + b.k - 1;
    const size_t a_tile = (size_t)C * xr1 * 4, t1 = (size_t)C * xr2 * 4, stage = (size_t)4 * C * C;
    const size_t misc = (17 + 2 * RB_MAX_STAGES) * 8 + 16 + 2 * C * 4 + 256;
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
    e = cudaFuncSetAttribute(rb_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(rb_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
// x planes (act applied by the producer) -> x' planes.  Returns 1 (launches), < 0 on error.
// `tiles` / `ntiles`: the super-tile table of this segmentation for ov = rb_ov(b) rows per tile (rb_tiles_kernel; ntiles is computed
// on the host from the same lengths).
inline int rb_ov(const RbWeights& b) { return 256 - (b.k - 1); }
inline int rb_pair_launch(int C, const RbWeights& a, const RbWeights& b, const Planes& in, const Planes& out, Seg seg, const int2* tiles, int ntiles,
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
    p.tmem_cols = 512;
    p.flags = flags;
    static const int env_dbg = getenv("STTS_RB_DBG") ? atoi(getenv("STTS_RB_DBG")) : 0;
    p.dbg = env_dbg;
    const long long W = ntiles;
    if (W <= 0 || !tiles) return -1;
    p.tiles = tiles; p.work_items = ntiles;
    alignas(64) CUtensorMap imap;
    if (!rb_make_map(&imap, in, p.xr1)) return -1;
    static const int env_tile_tma = getenv("STTS_TILE_TMA") ? atoi(getenv("STTS_TILE_TMA")) : 0;
    p.inp = in; p.bulk_in = env_tile_tma ? 0 : 1;
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
    const bool do_trace = env_trace && a.k == env_trace && a.dil == (a.k == 3 ? 1 : (a.k == 7 ? 3 : 5)) && trace_left[C == 64] > 0 && ntiles > 296;
    if (do_trace && !trace_buf) cudaMalloc(&trace_buf, 6 * 1024 * 8);
    if (do_trace) cudaMemsetAsync(trace_buf, 0, 6 * 1024 * 8, stream);
    p.trace = do_trace ? trace_buf : nullptr;
    if (C == 32) rb_pair_kernel<32><<<ctas, RB_THREADS, pl.smem, stream>>>(p, imap);
    else if (C == 16) rb_pair_kernel<16><<<ctas, RB_THREADS, pl.smem, stream>>>(p, imap);
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
        for (int r = 4; r < 6; ++r) {      // per-tap issue timeline of tile 1: (before elect, after syncwarp) pairs
            fprintf(stderr, " taps%d:", r - 4);
            for (int i = 0; i < 200; ++i) if (h[r * 1024 + i]) fprintf(stderr, " %lld", h[r * 1024 + i] - t0);
            fprintf(stderr, "\n");
        }
    }
    return 1;
}

}  // namespace stts
