// tools/mma_bench2.cu — round-2 tcgen05.mma micro-benchmark: cycles per MMA instruction (M = 128, K = 16, kind::f16)
// as a function of N and of WHERE the A operand lives (shared memory vs tensor memory), plus the two
// issue patterns the fused kernels use.  Model under test: an smem-A MMA costs max(N/2, (4096 + 32*N)/128)
// cycles (math vs shared-memory operand fetch at 128 B/clk), a TMEM-A MMA max(N/2, 32*N/128).
// GPU box tool; not part of the product.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_bench2 tools/mma_bench2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../summertts_b200/csrc/conv_tc.cuh"
using namespace stts;

struct Cfg { int N; int mode; };   // mode 0: smem A single stream; 1: merged pair (A_hi x 2N, A_lo x N); 2: TMEM A; 3: smem A, 3 MMAs (hi*hi, lo*hi, hi*lo)

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128) bench(const Cfg* cfgs, int ncfg, long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // fp16 1.0
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        uint32_t ph = 0;
        for (int c = 0; c < ncfg; ++c) {
            const Cfg cf = cfgs[c];
            const int N = cf.N;
            const uint32_t idN = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t id2N = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_s = smem_u32(sm), b_s = smem_u32(sm) + 96 * 1024;
            const uint32_t a_lbo = 160 * 16;   // A tile of 160 rows (halo), K halves 2560 B apart
            const uint64_t da = tc_desc(a_s, a_lbo, 128), db = tc_desc(b_s, (uint32_t)N * 16, 128), db2 = tc_desc(b_s, (uint32_t)N * 32, 128);
            const uint32_t a_lo_off = (40 * 1024) >> 4, b_lo_off = (16 * 1024) >> 4;
            for (int r = 0; r < 8; ++r) tc_mma_f16(tmem, da, db, idN, 1);
            tc_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
            const long long t0 = clock64();
            int issued = 0;
            if (cf.mode == 0) {
                for (int r = 0; r < reps; ++r) {
                    const uint64_t a = da + (uint64_t)((r & 3) * 2 * (a_lbo >> 4)) + (uint64_t)((r >> 2) & 7);   // moving K-step and tap shift
                    tc_mma_f16(tmem, a, db + (uint64_t)((r & 3) * 2 * N), idN, 1);
                    ++issued;
                }
            } else if (cf.mode == 1) {
                for (int r = 0; r < reps / 2; ++r) {
                    const uint64_t a = da + (uint64_t)((r & 3) * 2 * (a_lbo >> 4)) + (uint64_t)((r >> 2) & 7);
                    tc_mma_f16(tmem, a, db2 + (uint64_t)((r & 3) * 4 * N), id2N, 1);            // hi * [hi | lo]   (N' = 2N)
                    tc_mma_f16(tmem + 256, a + a_lo_off, db + (uint64_t)((r & 3) * 2 * N), idN, 1);   // lo * hi
                    issued += 2;
                }
            } else if (cf.mode == 2) {
                for (int r = 0; r < reps; ++r) {
                    mma_ts(tmem, tmem + 384 + (uint32_t)((r & 7) * 8), db + (uint64_t)((r & 3) * 2 * N), idN, 1);   // A: 128 lanes x 8 columns (16 halves)
                    ++issued;
                }
            } else {
                for (int r = 0; r < reps / 3; ++r) {
                    const uint64_t a = da + (uint64_t)((r & 3) * 2 * (a_lbo >> 4)) + (uint64_t)((r >> 2) & 7);
                    const uint64_t b = db + (uint64_t)((r & 3) * 2 * N);
                    tc_mma_f16(tmem, a, b, idN, 1);
                    tc_mma_f16(tmem + 256, a + a_lo_off, b, idN, 1);
                    tc_mma_f16(tmem + 256, a, b + b_lo_off, idN, 1);
                    issued += 3;
                }
            }
            tc_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
            const long long t1 = clock64();
            if (blockIdx.x == 0) { out[2 * c] = t1 - t0; out[2 * c + 1] = issued; }
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    const int reps = 240;
    std::vector<Cfg> c;
    for (int N : {16, 32, 64, 96, 128, 192, 256}) c.push_back({N, 0});
    for (int N : {32, 64, 96, 128}) c.push_back({N, 1});
    for (int N : {32, 64, 128, 192, 256}) c.push_back({N, 2});
    for (int N : {32, 64, 128, 192}) c.push_back({N, 3});
    static const char* names[4] = {"smem A, one stream        ", "merged: hi x [hi|lo] (2N) + lo x hi (N)", "TMEM A, one stream        ", "smem A, 3 MMAs per step   "};
    Cfg* d; long long* o;
    cudaMalloc(&d, c.size() * sizeof(Cfg)); cudaMalloc(&o, c.size() * 16);
    cudaMemcpy(d, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int grid : {1, 148}) {
        bench<<<grid, 128, 160 * 1024>>>(d, (int)c.size(), o, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(c.size() * 2);
        cudaMemcpy(h.data(), o, c.size() * 16, cudaMemcpyDeviceToHost);
        printf("grid=%d\n", grid);
        for (size_t i = 0; i < c.size(); ++i) {
            const double per = (double)h[2 * i] / (double)h[2 * i + 1];
            const int N = c[i].N;
            // algorithmic 128 x N x 16 MACs per K-step: mode 0/2 one MMA, mode 1 two MMAs, mode 3 three MMAs per algorithmic step
            const double per_step = per * (c[i].mode == 1 ? 2 : (c[i].mode == 3 ? 3 : 1));
            printf("  N=%3d %s : %7.1f cycles/MMA, %7.1f cycles per algorithmic K-step (math floor %d)\n", N, names[c[i].mode], per, per_step, N / 2);
        }
    }
    return 0;
}
