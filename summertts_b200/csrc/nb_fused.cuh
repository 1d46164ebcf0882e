// nb_fused.cuh — one ResBlock1 pair fused on the CUDA cores for NARROW stages (C = 4 / 8 channels: the 128x / 256x frame-rate
// stages of Generator_hifiGan, src/models/Generator_hifigan.cpp:154-173):
//
//     x' = x + conv2( leaky_0.1( conv1_dil( leaky_0.1(x) ) ) )                       (ResBlock1::forward, ResBlock1.cpp:55-69)
//
// At 4-8 channels a row carries 16-32 bytes and 2*k*C*C = 350-1400 MACs: the work is per-row, HBM-bound integer-width traffic, and
// a tensor-core tile pays its fixed per-tile epilogue for almost no math (rb_fused.cuh at C = 16 padding: slower than this).  So:
// one CTA stages leaky(x) for TR rows + halos in shared memory (coalesced float4 loads, zero outside the utterance), every thread
// computes whole rows of conv1 into a second shared tile (zero outside the utterance: conv2 pads ITS input), then whole rows of
// conv2 + bias + residual, and writes fp32 rows — 8 bytes of HBM per element per pair instead of 24, one launch instead of two.
// The weights of both convs travel as a __grid_constant__ kernel parameter (<= 5.7 KB): every FFMA takes its weight straight
// from the constant bank with a compile-time offset (taps, channels fully unrolled) — no weight loads at all in the inner loops.
// MRF (Generator_hifigan.cpp:160-171: xs = rb0; xs += rb1; xs += rb2; x = xs / 3) is folded into the last pair's store.
#pragma once
#include "kernels.cuh"

namespace stts {

constexpr int NB_THREADS = 256;
enum { NB_STORE = 0, NB_ACCUM = 1, NB_ACCUM_DIV = 2 };

template <int C, int K>
struct NbW {                       // both convs of the pair, [tap][ci][co]
    float w1[K * C * C], w2[K * C * C], b1[C], b2[C];
};
struct NbP {
    const float* x; float* y;      // [rows][C], row stride C
    const float* acc;              // NB_ACCUM / NB_ACCUM_DIV: y = (acc + x') [/ div]   (may alias y)
    Seg seg;
    int d1;                        // dilation of conv1 (conv2: 1)
    int tr;                        // output rows per CTA
    int out_mode; float div;
};

template <int C, int K>
__global__ void __launch_bounds__(NB_THREADS) nb_pair_kernel(const NbP p, const __grid_constant__ NbW<C, K> W) {
    extern __shared__ __align__(16) float nsm[];
    const int u = blockIdx.y;
    const int len = seg_len(p.seg, u);
    const int t0 = blockIdx.x * p.tr;
    if (t0 >= len) return;
    const int seg0 = seg_start(p.seg, u);
    const int h2 = (K - 1) / 2, h1 = h2 * p.d1, H = h1 + h2;
    const int nx = p.tr + 2 * H, nt = p.tr + 2 * h2;
    float* xs = nsm;                  // [nx][C]  leaky(x), rows t0 - H ...
    float* ts = nsm + (size_t)nx * C; // [nt][C]  leaky(conv1 + b1), rows t0 - h2 ...
    // ---- stage leaky(x) -----------------------------------------------------------------------
    {
        const int nvec = nx * C / 4;
        const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)seg0 * C);
        for (int i = threadIdx.x; i < nvec; i += NB_THREADS) {
            const int row = (i * 4) / C;              // C = 4: one float4 per row; C = 8: two
            const int tl = t0 - H + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tl >= 0 && tl < len) {
                v = __ldg(src + ((size_t)tl * C + (i * 4) % C) / 4);
                v.x = v.x < 0.f ? v.x * 0.1f : v.x; v.y = v.y < 0.f ? v.y * 0.1f : v.y;
                v.z = v.z < 0.f ? v.z * 0.1f : v.z; v.w = v.w < 0.f ? v.w * 0.1f : v.w;
            }
            reinterpret_cast<float4*>(xs)[i] = v;
        }
    }
    __syncthreads();
    // ---- conv1 (dilated) + bias + leaky -> ts, zero outside the utterance ---------------------
    for (int i = threadIdx.x; i < nt; i += NB_THREADS) {
        const int tl = t0 - h2 + i;
        float a[C];
#pragma unroll
        for (int co = 0; co < C; ++co) a[co] = W.b1[co];
        const float* xr = xs + (size_t)i * C;       // tap 0 reads x row (t0 - h2 + i) - h1 = xs row i
#pragma unroll
        for (int tap = 0; tap < K; ++tap) {
            float xv[C];
#pragma unroll
            for (int q = 0; q < C / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(xr + (size_t)tap * p.d1 * C + 4 * q);
                xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int ci = 0; ci < C; ++ci)
#pragma unroll
                for (int co = 0; co < C; ++co) a[co] = fmaf(xv[ci], W.w1[(tap * C + ci) * C + co], a[co]);
        }
        const bool valid = tl >= 0 && tl < len;
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            float4 v;
            v.x = a[4 * q]; v.y = a[4 * q + 1]; v.z = a[4 * q + 2]; v.w = a[4 * q + 3];
            v.x = v.x < 0.f ? v.x * 0.1f : v.x; v.y = v.y < 0.f ? v.y * 0.1f : v.y;
            v.z = v.z < 0.f ? v.z * 0.1f : v.z; v.w = v.w < 0.f ? v.w * 0.1f : v.w;
            if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(ts + (size_t)i * C + 4 * q) = v;
        }
    }
    __syncthreads();
    // ---- conv2 + bias + residual -> y ----------------------------------------------------------
    for (int r = threadIdx.x; r < p.tr; r += NB_THREADS) {
        const int tl = t0 + r;
        if (tl >= len) break;
        float a[C];
#pragma unroll
        for (int co = 0; co < C; ++co) a[co] = W.b2[co];
        const float* tr_ = ts + (size_t)r * C;      // tap 0 reads t1 row (t0 + r) - h2 = ts row r
#pragma unroll
        for (int tap = 0; tap < K; ++tap) {
            float tv[C];
#pragma unroll
            for (int q = 0; q < C / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(tr_ + (size_t)tap * C + 4 * q);
                tv[4 * q] = v.x; tv[4 * q + 1] = v.y; tv[4 * q + 2] = v.z; tv[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int ci = 0; ci < C; ++ci)
#pragma unroll
                for (int co = 0; co < C; ++co) a[co] = fmaf(tv[ci], W.w2[(tap * C + ci) * C + co], a[co]);
        }
        const size_t row = (size_t)(seg0 + tl) * C;
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + row) + q);      // raw x (L1/L2 hit): exact residual
            float4 v;
            v.x = a[4 * q] + xv.x; v.y = a[4 * q + 1] + xv.y; v.z = a[4 * q + 2] + xv.z; v.w = a[4 * q + 3] + xv.w;
            if (p.out_mode != NB_STORE) {
                const float4 o = *(reinterpret_cast<const float4*>(p.acc + row) + q);
                v.x = o.x + v.x; v.y = o.y + v.y; v.z = o.z + v.z; v.w = o.w + v.w;
                if (p.out_mode == NB_ACCUM_DIV) { v.x /= p.div; v.y /= p.div; v.z /= p.div; v.w /= p.div; }
            }
            *(reinterpret_cast<float4*>(p.y + row) + q) = v;
        }
    }
}

inline int nb_tile_rows(int C) { return C == 4 ? 2048 : 1024; }
inline size_t nb_smem(int C, int K, int d1) {
    const int h2 = (K - 1) / 2, H = h2 * d1 + h2, tr = nb_tile_rows(C);
    return ((size_t)(tr + 2 * H) + (size_t)(tr + 2 * h2)) * C * 4;
}
template <int C, int K>
inline cudaError_t nb_setup_one() { return cudaFuncSetAttribute(nb_pair_kernel<C, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); }
inline cudaError_t nb_device_setup() {
    cudaError_t e;
    if ((e = nb_setup_one<4, 3>()) != cudaSuccess) return e;
    if ((e = nb_setup_one<4, 7>()) != cudaSuccess) return e;
    if ((e = nb_setup_one<4, 11>()) != cudaSuccess) return e;
    if ((e = nb_setup_one<8, 3>()) != cudaSuccess) return e;
    if ((e = nb_setup_one<8, 7>()) != cudaSuccess) return e;
    return nb_setup_one<8, 11>();
}
inline bool nb_supported(int C, int k1, int d1, int pad1, int k2, int d2, int pad2) {
    if (C != 4 && C != 8) return false;
    if (k1 != k2 || (k1 != 3 && k1 != 7 && k1 != 11) || d2 != 1) return false;
    if (2 * pad1 != (k1 - 1) * d1 || 2 * pad2 != k2 - 1 || d1 < 1 || d1 > 8) return false;
    return nb_smem(C, k1, d1) <= 100 * 1024;
}
template <int C, int K>
inline void nb_launch_t(const NbP& p, const float* hw1, const float* hb1, const float* hw2, const float* hb2, int CoutW, int nseg, int maxlen,
                        cudaStream_t stream) {
    NbW<C, K> W;
    for (int t = 0; t < K; ++t)
        for (int ci = 0; ci < C; ++ci)
            for (int co = 0; co < C; ++co) {
                W.w1[(t * C + ci) * C + co] = hw1[((size_t)t * C + ci) * CoutW + co];
                W.w2[(t * C + ci) * C + co] = hw2[((size_t)t * C + ci) * CoutW + co];
            }
    for (int co = 0; co < C; ++co) { W.b1[co] = hb1 ? hb1[co] : 0.f; W.b2[co] = hb2 ? hb2[co] : 0.f; }
    dim3 g((maxlen + p.tr - 1) / p.tr, nseg);
    nb_pair_kernel<C, K><<<g, NB_THREADS, nb_smem(C, K, p.d1), stream>>>(p, W);
}
// host weights: [k][C][CoutW] fp32 (the layout make_conv builds) + bias[C] (or null)
inline int nb_pair_launch(int C, int k, NbP p, const float* hw1, const float* hb1, const float* hw2, const float* hb2, int CoutW, int nseg, int maxlen,
                          cudaStream_t stream) {
    p.tr = nb_tile_rows(C);
#define NB_CASE(CC, KK) if (C == CC && k == KK) { nb_launch_t<CC, KK>(p, hw1, hb1, hw2, hb2, CoutW, nseg, maxlen, stream); return 1; }
    NB_CASE(4, 3) NB_CASE(4, 7) NB_CASE(4, 11) NB_CASE(8, 3) NB_CASE(8, 7) NB_CASE(8, 11)
#undef NB_CASE
    return -1;
}

}  // namespace stts
