// tools/mma_bench3.cu — round-2 micro-benchmark, part 2 (GPU box tool; not part of the product):
//   A. does tcgen05.mma throughput for small N scale with the number of ISSUING warps?  (mma_bench2: one
//      issuing thread tops out at ~50-57 cycles per M=128,K=16 MMA for every N <= 96)
//   B. tcgen05.ld (TMEM -> registers) throughput with 4 / 8 warps
//   C. A and B together (promotion traffic next to the MMA stream)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/mma_bench3 tools/mma_bench3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../summertts_b200/csrc/conv_tc.cuh"
using namespace stts;

struct Cfg { int test; int N; int issuers; int ldwarps; };

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

__global__ void __launch_bounds__(384) bench(const Cfg* cfgs, int ncfg, long long* out, int reps, float* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    __shared__ long long t_start;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // fp16 1.0
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    uint32_t myph = 0;   // phase of this warp's own commit barrier (advances only when the warp issues)
    float acc = 0.f;
    for (int c = 0; c < ncfg; ++c) {
        const Cfg cf = cfgs[c];
        const int N = cf.N;
        __syncthreads();
        if (tid == 0) t_start = clock64();
        __syncthreads();
        // ---- issuing warps: warps 8..8+issuers-1 -------------------------------------------------
        if (warp >= 8 && warp < 8 + cf.issuers) {
            const int w = warp - 8;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_s = smem_u32(sm) + (uint32_t)w * 20 * 1024, b_s = smem_u32(sm) + 96 * 1024 + (uint32_t)w * 8 * 1024;
            const uint32_t a_lbo = 136 * 16;
            const uint64_t da = tc_desc(a_s, a_lbo, 128), db = tc_desc(b_s, (uint32_t)N * 16, 128);
            const uint32_t d_t = tmem + (uint32_t)w * 128;   // own accumulator (N <= 128)
            for (int r = 0; r < reps; ++r) {
                const uint64_t a = da + (uint32_t)((r & 3) * 2 * (a_lbo >> 4)) + (uint32_t)((r >> 2) & 7);
                const uint64_t b = db + (uint32_t)((r & 1) * 2 * N);
                if (elect_one()) tc_mma_f16(d_t, a, b, idesc, 1);
                __syncwarp();
            }
            if (elect_one()) tc_commit(&bars[w]);
            __syncwarp();
            mbar_wait_warp(&bars[w], myph);
            myph ^= 1;
            tc_fence_after();
        }
        // ---- TMEM load warps: warps 0..ldwarps-1 (lane quarter = warp % 4) -------------------------
        if (warp < cf.ldwarps) {
            const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
            const int lreps = cf.test == 0 ? 0 : reps;
            for (int r = 0; r < lreps; ++r) {
                uint32_t v[32];
                tc_ld32_nowait(tl + (uint32_t)((r & 1) * 32), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j += 8) acc += __uint_as_float(v[j]);
            }
        }
        __syncthreads();
        if (tid == 0 && blockIdx.x == 0) out[c] = clock64() - t_start;
    }
    if (acc == 123.456f) sink[tid] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    const int reps = 512;
    std::vector<Cfg> c;
    for (int N : {32, 64, 128})
        for (int iss : {1, 2, 4}) c.push_back({0, N, iss, 0});          // A: issue scaling
    c.push_back({1, 32, 0, 4});                                             // B: loads only, 4 warps (one per lane quarter)
    c.push_back({1, 32, 0, 8});                                             // B: 8 warps (two per lane quarter)
    for (int N : {32, 64}) { c.push_back({2, N, 2, 4}); c.push_back({2, N, 2, 8}); }   // C: 2 issuers + loads
    Cfg* d; long long* o; float* sink;
    cudaMalloc(&d, c.size() * sizeof(Cfg)); cudaMalloc(&o, c.size() * 8); cudaMalloc(&sink, 4096);
    cudaMemcpy(d, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int grid : {1, 148}) {
        bench<<<grid, 384, 160 * 1024>>>(d, (int)c.size(), o, reps, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(c.size());
        cudaMemcpy(h.data(), o, c.size() * 8, cudaMemcpyDeviceToHost);
        printf("grid=%d (reps=%d, times include ~100-200 cycles of barrier overhead)\n", grid, reps);
        for (size_t i = 0; i < c.size(); ++i) {
            const Cfg& f = c[i];
            const double cyc = (double)h[i];
            if (f.test == 0)
                printf("  A  N=%3d issuers=%d : %8.0f cycles, %6.1f cycles per MMA per issuer, %6.1f cycles per MMA aggregate (math floor %d)\n",
                       f.N, f.issuers, cyc, cyc / reps, cyc / (reps * f.issuers), f.N / 2);
            else if (f.test == 1)
                printf("  B  ld warps=%d    : %8.0f cycles, %6.1f B/clk aggregate (%d x 4 KB loads per warp)\n", f.ldwarps, cyc,
                       (double)f.ldwarps * reps * 4096.0 / cyc, reps);
            else
                printf("  C  N=%3d issuers=%d ld warps=%d : %8.0f cycles -> %6.1f cycles per MMA aggregate, %6.1f B/clk of tcgen05.ld\n", f.N,
                       f.issuers, f.ldwarps, cyc, cyc / (reps * f.issuers), (double)f.ldwarps * reps * 4096.0 / cyc);
        }
    }
    return 0;
}
