// Microbenchmark: issue cost (cycles per instruction, one warp, nothing else running on the SM) of the
// synchronisation instructions on the MMA issuer's critical path.  Build like tmem_bw.cu.
#include <cstdio>
#include <cstdlib>
#include "../../blamm_b200/csrc/filter_tc.cuh"
using namespace b200;

__global__ void __launch_bounds__(128, 1) issue_cost_kernel(unsigned long long* out, int n_mma)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) unsigned long long bar_s[8];
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (uint32_t i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    const uint32_t bar = smem_u32(&bar_s[0]);
    if (threadIdx.x == 0) { for (int q = 0; q < 8; q++) mbar_init(bar + 8 * q, 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        constexpr int R = 64;
        unsigned long long t[12];
        // completed phase 0 on barrier 1 so try_wait(parity 0) succeeds at once
        if (threadIdx.x == 0) mbar_arrive(bar + 8);
        __syncwarp();
        t[0] = clock64();
        for (int i = 0; i < R; i++) mbar_wait(bar + 8, 0, nullptr);
        t[1] = clock64();
        for (int i = 0; i < R; i++) tc_fence_after();
        t[2] = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < R; i++) acc += elect_one();
        t[3] = clock64();
        for (int i = 0; i < R; i++) __syncwarp();
        t[4] = clock64();
        for (int n : {32, 128, 256}) {
            const uint32_t idesc = (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t ad = umma_desc(smem_u32(smem), 32, 128), bd = umma_desc(smem_u32(smem) + 4096, 128, 256);
            uint32_t alo = (uint32_t)ad, ahi = (uint32_t)(ad >> 32), blo = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
            unsigned long long a = clock64();
            if (elect_one()) {
#pragma unroll 1
                for (int i = 0; i < n_mma; i++) { umma_f16_lohi(tmem, alo, ahi, blo, bhi, idesc, 1u); alo += 4; blo += 16; if (i % 8 == 7) { alo -= 32; blo -= 128; } }
            }
            __syncwarp();
            unsigned long long b = clock64();
            if (elect_one()) umma_commit(bar);
            __syncwarp();
            mbar_wait(bar, n == 32 ? 0 : (n == 128 ? 1 : 0), nullptr);
            unsigned long long c = clock64();
            const int k = n == 32 ? 5 : (n == 128 ? 7 : 9);
            t[k] = b - a; t[k + 1] = c - a;
        }
        unsigned long long a = clock64();
        if (elect_one()) {
#pragma unroll 1
            for (int i = 0; i < R; i++) umma_commit(bar + 16 + 8 * (i & 3));
        }
        __syncwarp();
        t[11] = clock64() - a;
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            out[0] = (t[1] - t[0]); out[1] = (t[2] - t[1]); out[2] = (t[3] - t[2]); out[3] = (t[4] - t[3]);
            for (int i = 5; i < 12; i++) out[i] = t[i];
            out[4] = acc;
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main()
{
    unsigned long long* d_out; cudaMalloc(&d_out, 16 * 8);
    cudaFuncSetAttribute(issue_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int n_mma : {8, 64}) {
        for (int rep = 0; rep < 2; rep++) issue_cost_kernel<<<148, 128, 65536>>>(d_out, n_mma);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[16]; cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        printf("per instruction (64 back to back): mbarrier.try_wait(hit) %.1f | tcgen05.fence::after %.1f | elect.sync %.1f | syncwarp %.1f | tcgen05.commit %.1f\n",
               h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[11] / 64.0);
        printf("  %d MMAs back to back (M=128, K=16): N=32 issue %.1f cyc/MMA, until complete %.1f | N=128 issue %.1f, complete %.1f | N=256 issue %.1f, complete %.1f\n",
               n_mma, (double)h[5] / n_mma, (double)h[6] / n_mma, (double)h[7] / n_mma, (double)h[8] / n_mma, (double)h[9] / n_mma, (double)h[10] / n_mma);
    }
    return 0;
}
