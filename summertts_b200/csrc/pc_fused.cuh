// pc_fused.cuh — wide-channel Conv1d on planes, activation tile resident per row tile, weights streamed (sm_100a, tcgen05 + TMEM + bulk copies):
// the two convs of a WaveNet layer (WN::forward, src/modules/WN.cpp:100-149; nn_conv1d.cpp:118-199)
//
//   EPI_GATE   acts = tanh(a[:, :H] + g) * sigmoid(a[:, H:] + g),  a = conv_k5(h)              (in_layer + fused_add_tanh_sigmoid_multiply,
//                                                                                               WN.cpp:85-98,118-126)
//   EPI_RS     rs = conv_1x1(acts);  h += rs[:, :H];  skip (+)= rs[:, H:]   (last layer: skip += rs)   (WN.cpp:128-146)
//
// Like rb_fused.cuh, every tensor is split-fp16 PLANES in HBM (8 x = hi + lo, [C/8][padded row][8], hi then lo):
//   * a work item = (128-row tile of one utterance, PAIR of 64-column output chunks); items are tile-major and every CTA owns a
//     CONTIGUOUS range of them, so the activation tile [2 Cin/8][XR][16 B] (one 1-D bulk copy per (plane, group) column) is
//     loaded once per tile and feeds all the chunk pairs of that tile that fall into the range.
//   * each chunk of the pair owns a TMEM slot of TWO accumulator buffers (128 columns each: hi*hi | hi*lo + lo*hi).  Promotion
//     units (8 K-steps = 2 weight stages of 64 input channels) alternate between the buffers and between two issuer threads:
//     the drain (tcgen05.ld -> fp32 registers) of unit u overlaps the MMAs of unit u + 1 of the same chunk AND the other chunk's.
//     The correction product A_lo x W_hi has the scale of the hi*lo half and accumulates onto it (same issuer thread, program order).
//   * weights stream from L2 through a bulk-copy ring in 16 KB stages [64 ch / 8][hi rows 64 | lo rows 64][8 halves].
//   * EPI_RS reads-modifies-writes the residual streams IN SHARED MEMORY: the producer loads the chunk's h / skip planes tile
//     into the staging buffer, the epilogue adds the conv result in place (x8 domain), bulk async stores write it back.
//   * EPI_GATE stores the 32 acts channels of a chunk straight from registers (lane = row: each warp store is 512 contiguous bytes).
// Arithmetic modes as in rb_fused.cuh (0: merged split-fp16 + promotion, fp32-accurate; 1: one fp16 MMA per K-step).
#pragma once
#include "rb_fused.cuh"

namespace stts {

constexpr int PC_THREADS = 704;     // warps 0-15: epilogue sets (slot x column half x lane quarter); 16-19: issuers (slot x accumulator buffer); 20: weight producer; 21: tile producer
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
    int nclu_items;                  // cluster work items: npairs * ceil(ntiles / ncta), tile-major (w = tile group * npairs + pair)
    int per;                         // items per cluster: cluster c owns the contiguous range [c * per, (c + 1) * per)
    Planes inp; int bulk_in;         // input planes; 1: tiles by 1-D bulk copies (planes_tile_g2s), 0: tensor-map boxes (STTS_TILE_TMA=1)
    int dbg;                         // timing experiments (STTS_PC_DBG; results are garbage): 1 = weight stages are not copied, 2 = MMAs are not issued, 4 = no epilogue math / stores
    long long* trace;                // STTS_TC_TRACE_BUILD: per-role cycle counters of CTA 0: [role][8]
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

// MMA issuer warp (slot, b) — the whole warp runs this, one elected lane issues: every promotion unit u = b (mod 2) of the slot's chunk goes into accumulator buffer b — main product
// A_hi x [W_hi | W_lo] (N = 128) and correction A_lo x W_hi (N = 64, accumulated onto the hi*lo half) from the SAME thread, so
// their order in the tensor pipe is the program order.  The two issuers of a slot alternate units: while buffer b drains, the
// other buffer's MMAs run.  Throughput mode: ONE N = 64 MMA per K-step, issuer b takes the stages st = b (mod 2), no promotion.
__device__ __forceinline__ void pc_issuer(const PcP& p, const int slot, const int b, const int w0, const int w1, const int rank, const uint32_t a_s0,
                                          const uint32_t w_s, const uint32_t tmem, uint64_t* a_full, uint64_t* a_empty, uint64_t* acc_full,
                                          uint64_t* acc_empty, uint64_t* b_full, uint64_t* b_empty) {
    const int mode = p.mode;
    const uint32_t idesc_m = (1u << 4) | ((uint32_t)((mode ? PC_NCH : 2 * PC_NCH) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_c = (1u << 4) | ((uint32_t)(PC_NCH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t b_lbo = 2 * PC_NCH * 16;
    constexpr uint32_t b_k16 = (2 * b_lbo) >> 4;
    const uint64_t b_desc0 = ((uint64_t)((b_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46) | (uint64_t)((w_s & 0x3FFFFu) >> 4);
    const uint32_t a_lbo = (uint32_t)p.xr * 16;
    const uint64_t a_bits = ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    const uint32_t a_k16 = (2 * a_lbo) >> 4;
    const uint32_t a_lo = (uint32_t)p.G * a_lbo;                    // the lo plane follows the hi plane in the tile
    const uint32_t d_m = tmem + (uint32_t)(slot * 256 + b * 128), d_c = d_m + 64u;
    uint64_t* full_bar = &acc_full[slot * 2 + b];
    uint64_t* empty_bar = &acc_empty[slot * 2 + b];
    const int nst = p.nst, nb = p.nb;
    const int SPU = max(1, p.usteps / 4);                           // weight stages per promotion unit (4 K-steps per stage)
    const int kcs = p.G / 8;                                        // 64-channel blocks per tap
    const uint16_t mc_mask = (uint16_t)((1u << p.ncta) - 1u);
    uint32_t af_par = 0, e_par = 1;
    int rs0 = slot; uint32_t rp0 = 0;                               // ring slot / phase of this chunk slot's next entry
    int prev_tg = -1;
#ifdef STTS_TC_TRACE_BUILD
    long long* tl_ = (p.trace && blockIdx.x == 0 && slot == 0) ? p.trace + 64 + b * 512 : nullptr; int titem = 0;
#define PC_TL(st, k) do { if (tl_ && titem < 8 && (threadIdx.x & 31) == 0) tl_[(titem * 15 + (st)) * 4 + (k)] = clock64(); } while (0)
    long long t_a = 0, t_e = 0, t_b = 0, t_all = clock64(), tt;
#define PC_T0() tt = clock64()
#define PC_T1(acc) acc += clock64() - tt
#else
#define PC_T0()
#define PC_T1(acc)
#define PC_TL(st, k)
#endif
#pragma unroll 1
    for (int w = w0; w < w1; ++w) {
        const int pair = w % p.npairs, tg = w / p.npairs, ti = tg * p.ncta + rank;
        const bool has_tile = ti < p.ntiles;                        // (a cluster's odd tail: only the weight-stage protocol runs)
        const bool active = has_tile && pair * 2 + slot < p.nchunks;
        if (has_tile && tg != prev_tg) { PC_T0(); mbar_wait_warp(a_full, af_par); af_par ^= 1; PC_T1(t_a); tc_fence_after(); }
        prev_tg = tg;
        // Both issuers of the slot walk EVERY stage of the slot's chunk and wait for its weights, whoever owns the stage: a ring
        // entry is refilled only after both have passed it, so neither can fall a whole mbarrier phase behind a ring slot
        // (a parity wait cannot tell fill n from fill n + 2).
        uint32_t acc = 0u;
        int tap = 0, kc = 0, uin = 0, ub = 0;                       // tap / 64-channel block of the stage; stage within its unit; unit parity
#pragma unroll 1
        for (int st = 0; st < nst; ++st) {
            const bool own = mode ? ((st & 1) == b) : (ub == b);
            const bool first = mode ? (st == b) : (uin == 0);
            const bool last = mode ? (st + 2 >= nst) : (uin == SPU - 1 || st == nst - 1);
            if (own && active && first) { PC_T0(); mbar_wait_warp(empty_bar, e_par); e_par ^= 1; PC_T1(t_e); tc_fence_after(); acc = 0u; }
            const int slot_r = rs0; const uint32_t ph_r = rp0;      // this slot's entries are every second ring entry, items back to back
            rs0 += 2; if (rs0 >= nb) { rs0 -= nb; rp0 ^= 1u; }
            PC_TL(st, 0);
            PC_T0(); mbar_wait_warp(&b_full[slot_r], ph_r); PC_T1(t_b);
            tc_fence_after();
            PC_TL(st, 1);
            if (elect_one()) {      // one election per stage: the MMAs and the commits that track them come from the same thread
            if (own && active && !(p.dbg & 2)) {
                const uint64_t da = a_bits | (uint64_t)(((a_s0 + (uint32_t)(tap * p.dil) * 16 + (uint32_t)(kc * 8) * a_lbo) & 0x3FFFFu) >> 4);
                const uint64_t dl = a_bits | (uint64_t)(((a_s0 + a_lo + (uint32_t)(tap * p.dil) * 16 + (uint32_t)(kc * 8) * a_lbo) & 0x3FFFFu) >> 4);
                const uint64_t db = b_desc0 + (uint32_t)slot_r * (uint32_t)(PC_STAGE >> 4);
                if (mode) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_m, da + (uint32_t)(ks * a_k16), db + (uint32_t)(ks * b_k16), idesc_m, ks == 0 ? acc : 1u);
                } else {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc_mma_f16(d_m, da + (uint32_t)(ks * a_k16), db + (uint32_t)(ks * b_k16), idesc_m, ks == 0 ? acc : 1u);
                        tc_mma_f16(d_c, dl + (uint32_t)(ks * a_k16), db + (uint32_t)(ks * b_k16), idesc_c, 1u);
                    }
                }
                acc = 1u;
            }
            if (p.ncta == 2) tc_commit_mc(&b_empty[slot_r], mc_mask); else tc_commit(&b_empty[slot_r]);
            if (own && active && last) tc_commit(full_bar);
            }
            PC_TL(st, 3);
            if (++kc == kcs) { kc = 0; ++tap; }
            if (++uin == SPU) { uin = 0; ub ^= 1; }
        }
#ifdef STTS_TC_TRACE_BUILD
        ++titem;
#endif
        if (has_tile && (w + 1 == w1 || (w + 1) / p.npairs != tg)) { if (elect_one()) tc_commit(a_empty); }   // this issuer's reads of the activation tile have retired
        __syncwarp();
    }
#ifdef STTS_TC_TRACE_BUILD
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { long long* o = p.trace + (b * 2 + slot) * 8; o[0] = clock64() - t_all; o[1] = t_a; o[2] = t_e; o[3] = t_b; }
#endif
}

template <int EPI>
__global__ void __launch_bounds__(PC_THREADS, 1) pc_kernel(const PcP p, const __grid_constant__ CUtensorMap imap, const __grid_constant__ CUtensorMap rmap0,
                                                           const __grid_constant__ CUtensorMap rmap1) {
    constexpr int SG = EPI == PC_EPI_GATE ? 4 : 8;          // 16-byte channel groups per plane of a chunk's output (32 / 64 channels)
    constexpr int STG = EPI == PC_EPI_GATE ? 0 : 2 * SG * 128 * 16;      // staging bytes per slot (RS: 32 KB; GATE stores straight from registers)
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
    uint64_t* acc_full = bars + 2;       // [slot][buffer]
    uint64_t* acc_empty = bars + 6;      // [slot][buffer]
    uint64_t* r_full = bars + 10;        // [2] residual tile landed in the slot's staging buffer (RS)
    uint64_t* s_free = bars + 12;        // [2] staging buffer free again (its bulk stores have read it)
    uint64_t* b_full = bars + 14;        // [PC_MAX_RING]
    uint64_t* b_empty = b_full + PC_MAX_RING;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + PC_MAX_RING);

    if (tid == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 4);
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        for (int i = 0; i < 2; ++i) { mbar_init(&r_full[i], 1); mbar_init(&s_free[i], 1); }
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
    const uint32_t tmem = *tmem_slot;     // slot s, buffer b @ s*256 + b*128: 64 columns hi*hi | 64 columns hi*lo + lo*hi
    if (p.ncta == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into them
    // contiguous range of (tile-major) cluster items: consecutive items share the activation tile, which is loaded once per tile
    const int crank = p.ncta == 2 ? (int)cluster_rank() : 0;
    const int w0 = min(p.nclu_items, (int)(blockIdx.x / p.ncta) * p.per), w1 = min(p.nclu_items, w0 + p.per);

    if (warp < 16) {
        // ================= promotion + epilogue: (slot, column half hf), one row x 32 columns per thread ==================
        const int wq = warp & 3, slot = (warp >> 2) & 1, hf = warp >> 3;
        const int tl = wq * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(slot * 256 + hf * 32);
        uint8_t* sbuf = stg + (size_t)slot * STG;
        uint32_t f_par = 0, rf_par = 0;      // f_par: bit b = parity of acc_full[slot][b]
        float amax = 0.f;
        float racc[32];
        const int SPU = max(1, p.usteps / 4);
        const int NU = p.mode ? 2 : (p.nst + SPU - 1) / SPU;
#ifdef STTS_TC_TRACE_BUILD
        long long* el_ = (p.trace && blockIdx.x == 0 && warp == 0 && lane == 0) ? p.trace + 64 + 1024 : nullptr; int eitem = 0;
#define PC_EL(un, k) do { if (el_ && eitem < 8) el_[(eitem * 9 + (un)) * 3 + (k)] = clock64(); } while (0)
#else
#define PC_EL(un, k)
#endif
#pragma unroll 1
        for (int w = w0; w < w1; ++w) {
            const int pair = w % p.npairs, ti = (w / p.npairs) * p.ncta + crank;
            const int chunk = pair * 2 + slot;
            const bool active = chunk < p.nchunks && ti < p.ntiles;
            if (!active) continue;                    // (the partner slot still runs; nothing of this slot's barriers is used)
            const RbTile it = rb_tile_at(p.seg, p.tiles, ti);
#pragma unroll 1
            for (int un = 0; un < NU; ++un) {
                const int b = un & 1;
                PC_EL(un, 0);
                mbar_wait_all(&acc_full[slot * 2 + b], (f_par >> b) & 1u); f_par ^= 1u << b;
                tc_fence_after();
                PC_EL(un, 1);
                const uint32_t tsrc = tbase + (uint32_t)(b * 128);
#pragma unroll
                for (int cb = 0; cb < 32; cb += 16) {
                    uint32_t v[16], x2[16];
                    tc_ld_nowait<16>(tsrc + cb, v);
                    if (!p.mode) tc_ld_nowait<16>(tsrc + 64 + cb, x2);
                    tc_ld_wait();
                    if (p.mode) {
                        if (un == 0) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) racc[cb + j] = __uint_as_float(v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) racc[cb + j] += __uint_as_float(v[j]);
                        }
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
                if (lane == 0) mbar_arrive(&acc_empty[slot * 2 + b]);
                PC_EL(un, 2);
            }
            PC_EL(8, 0);
            const int tr = it.t0 + tl;
            const bool valid = tr < it.len;
            const int n0 = chunk * PC_NCH + hf * 32;       // first output column of this thread
            if (EPI == PC_EPI_GATE && (p.dbg & 4)) {
            } else if (EPI == PC_EPI_GATE) {
                // acts planes straight from registers: lane = row, so every 16-byte store of a warp lands in 512 contiguous bytes
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
                if (valid) {                               // rows past the utterance stay untouched: the gap rows are zero, the next utterance is not ours
                    const int Gout = p.out0.C / 8;
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg) {
                        uint4 hi, lo;
                        split8_scaled(o + 8 * gg, hi, lo);
                        __half* dh = p.out0.base + ((size_t)(chunk * SG + hf * 2 + gg) * p.out0.rows_p + (size_t)(it.prow_u + tr)) * 8;
                        *reinterpret_cast<uint4*>(dh) = hi;
                        *reinterpret_cast<uint4*>(dh + (size_t)Gout * p.out0.rows_p * 8) = lo;
                    }
                }
            } else {
                const bool to1 = chunk >= p.split;
                const Planes& op = to1 ? p.out1 : p.out0;
                const int oc = to1 ? chunk - p.split : chunk;          // 64-channel chunk inside the destination stream
                const bool accin = to1 ? p.acc1 != 0 : p.acc0 != 0;
                // the producer hands the staging tile over once per chunk — with the stream's previous value in it when the chunk
                // accumulates, empty otherwise — and only after the previous chunk's bulk stores have read it (s_free).  Waiting
                // here in both cases keeps the epilogue from running a whole s_free phase ahead of the producer's parity test.
                mbar_wait_all(&r_full[slot], rf_par); rf_par ^= 1;
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
            PC_EL(8, 1);
#ifdef STTS_TC_TRACE_BUILD
            ++eitem;
#endif
        }
        if (EPI == PC_EPI_RS && tl < 2 * SG && hf == 0) bulk_wait_all0();
        if (amax > 65000.f && p.flags && !p.dbg) atomicOr(p.flags, 1u);
    } else if (warp < 20) {
        // the whole warp runs the issuer loop (warp-uniform control flow keeps the descriptors in uniform registers); one elected
        // lane issues the MMAs and commits
        pc_issuer(p, (warp - 16) & 1, (warp - 16) >> 1, w0, w1, crank, smem_u32(abuf), smem_u32(wst), tmem, a_full, a_empty, acc_full, acc_empty, b_full,
                  b_empty);
        __syncwarp();
    } else if (warp == 20) {
        // ================= weight producer: one thread, a tight blocking loop (a few instructions per 16 KB stage) ==============
        // (A single thread polling all three streams cooperatively paced the whole kernel: ~100 dependent instructions per stage,
        //  each ~10 cycles on an SM whose schedulers are shared with 16 polling epilogue warps — 1.2k cycles per stage, measured.)
        if (lane == 0) {
            const uint16_t mc_mask = (uint16_t)((1u << p.ncta) - 1u);
            int ws_ = 0; uint32_t wph = 1;
            int pair = w0 % p.npairs;
#ifdef STTS_TC_TRACE_BUILD
            int pfill = 0;
            if (p.trace && blockIdx.x == 0) p.trace[63] = clock64();
#endif
#pragma unroll 1
            for (int w = w0; w < w1; ++w) {
                // ring order: stage 0 of slot 0, stage 0 of slot 1, stage 1 of slot 0, ...; an inactive slot (odd chunk count) still gets
                // (ignored) bytes: the ring walk stays uniform
                const uint8_t* src0 = reinterpret_cast<const uint8_t*>(p.w) + (size_t)min(pair * 2, p.nchunks - 1) * p.nst * PC_STAGE;
                const uint8_t* src1 = reinterpret_cast<const uint8_t*>(p.w) + (size_t)min(pair * 2 + 1, p.nchunks - 1) * p.nst * PC_STAGE;
#pragma unroll 1
                for (int st = 0; st < p.nst; ++st) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const uint8_t* src = (q ? src1 : src0) + (size_t)st * PC_STAGE;
                        mbar_wait(&b_empty[ws_], wph);
#ifdef STTS_TC_TRACE_BUILD
                        if (p.trace && blockIdx.x == 0 && pfill < 256) p.trace[64 + 1024 + 256 + pfill] = clock64();
                        ++pfill;
#endif
                        if (p.dbg & 1) mbar_arrive(&b_full[ws_]);
                        else {
                            mbar_expect_tx(&b_full[ws_], PC_STAGE);
                            if (p.ncta == 2) {        // this CTA fetches its half of the stage and multicasts it into both CTAs (same offsets, same barrier)
                                const uint32_t half = PC_STAGE / 2;
                                bulk_g2s_mc(wst + (size_t)ws_ * PC_STAGE + crank * half, src + crank * half, half, &b_full[ws_], mc_mask);
                            } else bulk_g2s(wst + (size_t)ws_ * PC_STAGE, src, PC_STAGE, &b_full[ws_]);
                        }
                        if (++ws_ == p.nb) { ws_ = 0; wph ^= 1; }
                    }
                }
                if (++pair == p.npairs) pair = 0;
            }
        }
        __syncwarp();
    } else {
        // ================= tile producer (warp 21): activation tiles and, for RS, the residual tiles, in item order ============
        // Lane 0 waits (blocking) and posts the expected bytes; all lanes then issue the per-(plane, group) bulk copies of the tile.
        // Order per item: [first item of a tile: a_empty -> activation tile]  slot 0 residual  slot 1 residual.  Everything a wait
        // here depends on (the issuers' a_empty, the epilogue's s_free) needs only loads that come EARLIER in this sequence.
        uint32_t ae_par = 1, sf_par0 = 1, sf_par1 = 1;
        int pair = w0 % p.npairs, tg = w0 / p.npairs, prev_tg = -1;
#pragma unroll 1
        for (int w = w0; w < w1; ++w) {
            const int ti = tg * p.ncta + crank;
            const bool has_tile = ti < p.ntiles;       // (odd tail of a cluster: no tile for this CTA)
            if (has_tile) {
                const RbTile it = rb_tile_at(p.seg, p.tiles, ti);
                if (tg != prev_tg) {
                    if (lane == 0) { mbar_wait(a_empty, ae_par); mbar_expect_tx(a_full, a_tile); }
                    ae_par ^= 1;
                    __syncwarp();
                    const long long r0 = it.prow_u + it.t0 - p.padl;
                    if (p.bulk_in) {
                        for (int g = lane; g < 2 * p.G; g += 32) planes_tile_g2s(abuf + (size_t)g * p.xr * 16, p.inp, g, 1, r0, p.xr, a_full);
                    } else if (lane == 0) tma_load_3d(abuf, &imap, 0, (int)r0, 0, a_full);
                }
                if (EPI == PC_EPI_RS) {
#pragma unroll
                    for (int rslot = 0; rslot < 2; ++rslot) {
                        const int chunk = pair * 2 + rslot;
                        if (chunk >= p.nchunks) continue;
                        const bool to1 = chunk >= p.split;
                        const bool need = to1 ? p.acc1 != 0 : p.acc0 != 0;
                        const Planes& op = to1 ? p.out1 : p.out0;
                        const int oc = to1 ? chunk - p.split : chunk;
                        uint8_t* dst = stg + (size_t)rslot * STG;
                        // every chunk's epilogue frees the staging tile once (s_free) and waits for it once (r_full): the stream's
                        // previous value is loaded when the chunk accumulates, otherwise the tile is handed over empty
                        if (lane == 0) {
                            mbar_wait(&s_free[rslot], rslot ? sf_par1 : sf_par0);
                            if (need) mbar_expect_tx(&r_full[rslot], STG); else mbar_arrive(&r_full[rslot]);
                        }
                        if (rslot) sf_par1 ^= 1; else sf_par0 ^= 1;
                        __syncwarp();
                        if (need) {
                            const long long row = it.prow_u + it.t0;
                            if (p.bulk_in) {
                                if (lane < 2 * SG) {
                                    const int plane = lane / SG, g = lane - plane * SG;
                                    planes_tile_g2s(dst + (size_t)plane * (STG / 2) + (size_t)g * 128 * 16, op, plane * (op.C / 8) + oc * SG + g, 1, row, 128, &r_full[rslot]);
                                }
                            } else if (lane == 0) {
                                const CUtensorMap* rm = to1 ? &rmap1 : &rmap0;
                                tma_load_3d(dst, rm, 0, (int)row, oc * SG, &r_full[rslot]);
                                tma_load_3d(dst + STG / 2, rm, 0, (int)row, op.C / 8 + oc * SG, &r_full[rslot]);
                            }
                        }
                    }
                }
            }
            prev_tg = tg;
            if (++pair == p.npairs) { pair = 0; ++tg; }
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
    const size_t stg = epi == PC_EPI_GATE ? 0 : (size_t)2 * 32 * 1024;      // GATE stores straight from registers
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
    p.per = (p.nclu_items + ctas / p.ncta - 1) / (ctas / p.ncta);
    static const int env_tile_tma = getenv("STTS_TILE_TMA") ? atoi(getenv("STTS_TILE_TMA")) : 0;
    p.inp = in; p.bulk_in = env_tile_tma ? 0 : 1;
    static const int env_dbg = getenv("STTS_PC_DBG") ? atoi(getenv("STTS_PC_DBG")) : 0;
    p.dbg = epi == PC_EPI_GATE ? env_dbg : 0;
    static long long* trace_buf = nullptr;
    static const int env_trace = getenv("STTS_PC_TRACE") ? atoi(getenv("STTS_PC_TRACE")) : 0;
    static int trace_left[2] = {2, 2};
    const bool do_trace = env_trace && trace_left[epi] > 0 && ntiles > 100;
    if (do_trace && !trace_buf) cudaMalloc(&trace_buf, 2048 * 8);
    if (do_trace) cudaMemsetAsync(trace_buf, 0, 2048 * 8, stream);
    p.trace = do_trace ? trace_buf : nullptr;
    cudaError_t le;
    if (epi == PC_EPI_GATE) le = cudaLaunchKernelEx(&cfg, pc_kernel<PC_EPI_GATE>, p, imap, rmap0, rmap1);
    else le = cudaLaunchKernelEx(&cfg, pc_kernel<PC_EPI_RS>, p, imap, rmap0, rmap1);
    if (do_trace) {
        --trace_left[epi];
        static long long h[2048];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        if (const char* tf = getenv("STTS_PC_TRACE_FILE")) {
            char nm[512]; snprintf(nm, sizeof(nm), "%s.epi%d.%d", tf, epi, trace_left[epi]);
            if (FILE* f = fopen(nm, "w")) { for (int i = 0; i < 2048; ++i) fprintf(f, "%lld\n", h[i]); fclose(f); }
        }
        fprintf(stderr, "PCTRACE epi=%d mode=%d items/cta %d nb=%d:", epi, mode, p.per, p.nb);
        for (int r = 0; r < 4; ++r) fprintf(stderr, "  [buf%d slot%d total %lld a_full %lld empty %lld b_full %lld]", r >> 1, r & 1, h[r * 8], h[r * 8 + 1], h[r * 8 + 2], h[r * 8 + 3]);
        fprintf(stderr, "\n");
    }
    return le == cudaSuccess ? 1 : -5;
}

}  // namespace stts
