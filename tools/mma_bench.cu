// tools/mma_bench.cu — tcgen05.mma issue-rate microbenchmark (cycles per MMA instruction) for the
// operand layouts conv_tc.cuh uses or might use.  GPU box tool; not part of the product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../summertts_b200/csrc/conv_tc.cuh"
using namespace stts;

struct Cfg { int N; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; uint32_t a_shift; int layout; int two_acc; int real; };

__device__ __forceinline__ uint64_t desc_l(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128) bench(const Cfg* cfgs, int ncfg, long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar, bar_done[2], bar_sink;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // fp16 1.0
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1); mbar_init(&bar_sink, 1); mbar_arrive(&bar_done[0]); mbar_arrive(&bar_done[1]); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
            const uint32_t idesc = (1u << 4) | ((uint32_t)(cf.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_s = smem_u32(sm) + cf.a_shift, b_s = smem_u32(sm) + 96 * 1024;
            const uint64_t da = desc_l(a_s, cf.a_lbo, cf.a_sbo, cf.layout), db = desc_l(b_s, cf.b_lbo, cf.b_sbo, cf.layout);
            // warm-up
            for (int r = 0; r < 8; ++r) tc_mma_f16(tmem, da, db, idesc, 1);
            tc_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
            const long long t0 = clock64();
            if (cf.real) {
                // like conv_tc: per K-step (A_hi,B_hi)->main, (A_lo,B_hi)->corr, (A_hi,B_lo)->corr; operands move every step
                const uint32_t a_lo_off = (36 * 1024) >> 4, b_lo_off = (32 * 1024) >> 4;
                for (int r = 0; r < reps / 3; ++r) {
                    if (cf.real >= 2 && (r % cf.real) == 0) { mbar_wait(&bar_done[r & 1], 0); tc_fence_after(); }   // already-complete barrier
                    if (cf.real >= 2 && (r % cf.real) == cf.real - 1) tc_commit(&bar_sink);                       // commit nobody waits on
                    const uint64_t a = da + (uint64_t)((r & 7) * 2 * (cf.a_lbo >> 4)) + (uint64_t)((r >> 3) & 3);
                    const uint64_t b = db + (uint64_t)((r & 3) * 2 * (cf.b_lbo >> 4));
                    tc_mma_f16(tmem, a, b, idesc, 1);
                    tc_mma_f16(tmem + 256, a + a_lo_off, b, idesc, 1);
                    tc_mma_f16(tmem + 256, a, b + b_lo_off, idesc, 1);
                }
            } else
            for (int r = 0; r < reps; ++r) tc_mma_f16(tmem + ((cf.two_acc && (r & 1)) ? 256 : 0), da, db, idesc, 1);
            tc_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[c] = t1 - t0;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main(int argc, char** argv) {
    const int reps = 258;
    std::vector<Cfg> c;
    std::vector<const char*> names;
    auto add = [&](const char* n, Cfg f) { c.push_back(f); names.push_back(n); };
    // no-swizzle K-major: rows 16 B apart (SBO 128), K-halves LBO apart
    for (int N : {32, 64, 128, 192, 256}) {
        add("NONE  REAL pattern (3 MMAs/step, moving operands)", {N, 2112, 128, (uint32_t)N * 16, 128, 0, 0, 0, 1});
        add("NONE  REAL + wait/commit every 2 steps (6 MMAs)  ", {N, 2112, 128, (uint32_t)N * 16, 128, 0, 0, 0, 2});
        add("NONE  REAL + wait/commit every 4 steps (12 MMAs) ", {N, 2112, 128, (uint32_t)N * 16, 128, 0, 0, 0, 4});
        add("NONE  a_lbo=2112(132 rows) b_lbo=N*16        ", {N, 2112, 128, (uint32_t)N * 16, 128, 0, 0, 0, 0});
        add("NONE  a_lbo=2048(128 rows) b_lbo=N*16        ", {N, 2048, 128, (uint32_t)N * 16, 128, 0, 0, 0, 0});
        add("NONE  a_lbo=2080 b_lbo=N*16+32 (both offset) ", {N, 2080, 128, (uint32_t)N * 16 + 32, 128, 0, 0, 0, 0});
        add("NONE  a shifted by 3 rows                    ", {N, 2112, 128, (uint32_t)N * 16, 128, 48, 0, 0, 0});
        add("SW128 (SBO 1024)                             ", {N, 16, 1024, 16, 1024, 0, 2, 0, 0});
        add("SW128 alternating 2 accumulators             ", {N, 16, 1024, 16, 1024, 0, 2, 1, 0});
        add("NONE  alternating 2 accumulators             ", {N, 2112, 128, (uint32_t)N * 16, 128, 0, 0, 1, 0});
    }
    Cfg* d; long long* o;
    cudaMalloc(&d, c.size() * sizeof(Cfg)); cudaMalloc(&o, c.size() * 8);
    cudaMemcpy(d, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int grid : {1, 148}) {
        bench<<<grid, 128, 160 * 1024>>>(d, (int)c.size(), o, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(c.size());
        cudaMemcpy(h.data(), o, c.size() * 8, cudaMemcpyDeviceToHost);
        printf("grid=%d\n", grid);
        for (size_t i = 0; i < c.size(); ++i)
            printf("  N=%3d %s : %7.1f cycles/MMA  (floor %d)\n", c[i].N, names[i], (double)h[i] / reps, c[i].N / 2);
    }
    return 0;
}
