// tools/mma_bench4.cu — why does the fused kernel's issuer see ~300 cycles per MMA when mma_bench3 sees 40-50?
// One variable at a time on top of the 2-issuer N=64 / N=32 pattern: (P) 8 other warps polling an mbarrier (lane 0,
// try_wait loop; with and without nanosleep back-off), (L) A-tile LBO 2848 B + 5-row tap shifts, (C) a tcgen05.commit
// after every pair of MMAs, (V) descriptors rebuilt from values loaded from memory (vector registers -> R2UR).
// GPU box tool; not part of the product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../summertts_b200/csrc/conv_tc.cuh"
using namespace stts;

struct Cfg { int poll; int lbo_rows; int shift; int commit; int vec; };

__device__ __forceinline__ void mbar_wait_sleep(uint64_t* b, uint32_t parity, unsigned ns) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!done && ns) __nanosleep(ns);
    }
}

__global__ void __launch_bounds__(384) bench(const Cfg* cfgs, int ncfg, long long* out, int reps, const int* vals) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bars[8];
    __shared__ uint64_t never;
    __shared__ uint64_t sink;
    __shared__ uint32_t slot;
    __shared__ long long t_start;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&never, 1); mbar_init(&sink, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    uint32_t myph = 0, nvph = 0;
    for (int c = 0; c < ncfg; ++c) {
        const Cfg cf = cfgs[c];
        __syncthreads();
        if (tid == 0) t_start = clock64();
        __syncthreads();
        if (warp >= 8 && warp < 10) {
            const int w = warp - 8;
            const int N = w == 0 ? 64 : 32;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_lbo = (uint32_t)cf.lbo_rows * 16;
            uint32_t a_s = smem_u32(sm) + (w ? 4 * a_lbo : 0), b_s = smem_u32(sm) + 96 * 1024;
            if (cf.vec) { a_s += (uint32_t)vals[lane & 1]; b_s += (uint32_t)vals[(lane & 1) + 2]; }     // zeros, but loaded: vector registers
            const uint64_t a_bits = ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint64_t b_bits = ((uint64_t)((1024 >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint32_t d_t = tmem + (uint32_t)w * 128;
            for (int r = 0; r < reps; ++r) {
                const int tap = r % 11;
                const uint64_t da = a_bits | (uint64_t)(((a_s + (uint32_t)(tap * cf.shift) * 16) & 0x3FFFFu) >> 4);
                const uint64_t db = b_bits | (uint64_t)(((b_s + (uint32_t)(tap) * 4096) & 0x3FFFFu) >> 4);
                if (elect_one()) {
                    tc_mma_f16(d_t, da, db, idesc, 1);
                    tc_mma_f16(d_t, da + (uint32_t)((2 * a_lbo) >> 4), db + (uint32_t)(2048 >> 4), idesc, 1);
                    if (cf.commit) tc_commit(&sink);
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(&bars[w]);
            __syncwarp();
            mbar_wait_warp(&bars[w], myph);
            myph ^= 1;
            tc_fence_after();
            if (w == 0 && lane == 0) mbar_arrive(&never);      // release the pollers
        } else if (warp < 8 && cf.poll) {
            if (lane == 0) mbar_wait_sleep(&never, nvph, cf.poll == 2 ? 200u : 0u);
            __syncwarp();
        } else if (warp < 8) {
        }
        // every config completes `never` exactly once so the phases stay aligned
        if (!(cf.poll) ) { }
        __syncthreads();
        if (tid == 0 && blockIdx.x == 0) out[c] = clock64() - t_start;
        nvph ^= 1;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    const int reps = 440;
    std::vector<Cfg> c = {
        {0, 160, 1, 0, 0},   // base: aligned LBO, 1-row shifts
        {1, 160, 1, 0, 0},   // + 8 polling lanes (tight try_wait loop)
        {2, 160, 1, 0, 0},   // + 8 polling lanes with 200 ns back-off
        {0, 178, 5, 0, 0},   // LBO 2848 B, 5-row shifts
        {0, 160, 1, 1, 0},   // commit after every MMA pair
        {0, 160, 1, 0, 1},   // descriptors from loaded values
        {1, 178, 5, 1, 1},   // everything
        {2, 178, 5, 1, 1},   // everything, polite polling
    };
    static const char* names[] = {"base", "+poll tight", "+poll 200ns", "LBO 2848, shift 5", "+commit per pair", "+vector descriptors", "all (tight poll)", "all (200 ns poll)"};
    Cfg* d; long long* o; int* vals;
    cudaMalloc(&d, c.size() * sizeof(Cfg)); cudaMalloc(&o, c.size() * 8); cudaMalloc(&vals, 16); cudaMemset(vals, 0, 16);
    cudaMemcpy(d, c.data(), c.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int grid : {1, 148}) {
        bench<<<grid, 384, 160 * 1024>>>(d, (int)c.size(), o, reps, vals);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(c.size());
        cudaMemcpy(h.data(), o, c.size() * 8, cudaMemcpyDeviceToHost);
        printf("grid=%d: two issuers (N=64 | N=32), %d iterations x 2 MMAs each\n", grid, reps);
        for (size_t i = 0; i < c.size(); ++i)
            printf("  %-22s : %8lld cycles, %6.1f cycles per iteration (2 MMAs per issuer), i.e. %5.1f per MMA aggregate\n", names[i], h[i], (double)h[i] / reps, (double)h[i] / (reps * 4));
    }
    return 0;
}
