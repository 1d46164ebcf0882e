// tools/tma_bench.cu — how fast does an activation tile [2G groups][XR rows][16 B] reach shared memory?
//   (a) one 3-D tensor-map box {8 halves, XR rows, 2G groups}   (what conv_tc / rb_fused / pc_fused used: 16-byte inner extent)
//   (b) 2G one-dimensional bulk copies of XR*16 B               (each (plane, group) column of the planes layout is contiguous)
// All CTAs load concurrently from an L2-resident planes buffer; reports cycles per tile and B/clk/SM.  GPU box tool; not part of the product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../summertts_b200/csrc/pc_fused.cuh"
using namespace stts;

__global__ void __launch_bounds__(128) bench(const __grid_constant__ CUtensorMap map, const __half* base, long long rows_p, int G2, int xr, int mode, int reps,
                                             long long* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t par = 0;
        const uint32_t bytes = (uint32_t)G2 * xr * 16;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const long long row0 = ((long long)(blockIdx.x * 37 + r * 151) * 128) % (rows_p - 512);
            mbar_expect_tx(&bar, bytes);
            if (mode == 0) tma_load_3d(sm, &map, 0, (int)row0, 0, &bar);
            else
                for (int g = 0; g < G2; ++g) bulk_g2s(sm + (size_t)g * xr * 16, base + ((size_t)g * rows_p + row0) * 8, (uint32_t)xr * 16, &bar);
            mbar_wait(&bar, par); par ^= 1;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const int C = 192, G2 = 2 * C / 8, xr = 132;
    const long long rows_p = 50000;
    Planes pl; pl.C = C; pl.rows_p = rows_p;
    cudaMalloc(&pl.base, (size_t)G2 * rows_p * 16);
    cudaMemset(pl.base, 0, (size_t)G2 * rows_p * 16);
    alignas(64) CUtensorMap map;
    if (!pc_make_map(&map, pl, xr, G2)) { printf("map failed\n"); return 1; }
    long long* d_out; cudaMalloc(&d_out, 148 * 8);
    const size_t smem = (size_t)G2 * xr * 16 + 128;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 40;
    for (int ctas : {1, 148})
        for (int mode = 0; mode < 2; ++mode)
            for (int it = 0; it < 2; ++it) {
                bench<<<ctas, 128, smem>>>(map, pl.base, rows_p, G2, xr, mode, reps, d_out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                std::vector<long long> h(ctas);
                cudaMemcpy(h.data(), d_out, ctas * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; double avg = 0; for (auto v : h) { mx = std::max(mx, v); avg += v; } avg /= ctas;
                const double bytes = (double)G2 * xr * 16;
                printf("ctas %3d mode %s run %d: %.0f cycles/tile avg (max %.0f)  %.1f B/clk/SM  tile %.0f B\n", ctas, mode ? "bulk1d" : "tma3d ", it, avg / reps,
                       (double)mx / reps, bytes / (avg / reps), bytes);
            }
    return 0;
}
