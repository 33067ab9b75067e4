// Microbenchmark: tcgen05.ld throughput out of TMEM, alone and against a saturated tcgen05.mma stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tools/micro/tmem_bw.cu ; run on a B200.
// Question it answers: is the filter epilogue bound by TMEM read bandwidth or by latency?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../blamm_b200/csrc/filter_tc.cuh"
using namespace b200;

struct Cfg { int epi_warps; int pack; int mma; int lds_per_wait; int n_k; int iters; int shape; int n; int sbo; int commit; int fill; };

template <int X> __device__ __forceinline__ void ld_any(uint32_t taddr, int pack, uint32_t (&v)[32]);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(32 * 17, 1) tmem_bw_kernel(Cfg c, unsigned long long* out, uint32_t* sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) unsigned long long bar_s[8];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = c.fill ? ((i * 2654435761u) & 0x3C003C00u) : 0u;
    const uint32_t bar = smem_u32(&bar_s[0]);
    if (threadIdx.x == 0) { for (int q = 0; q < 8; q++) mbar_init(bar + 8 * q, 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    unsigned long long t0 = clock64(), t1 = t0;
    if (warp == 0) {
        if (c.mma) {
            // idesc: D fp16 (bits 4-5 = 0) or f32 (=1); A,B fp16 (0); N>>3 at bit 17; M>>4 at bit 24
            const uint32_t idesc = (((uint32_t)c.n >> 3) << 17) | ((128u >> 4) << 24) | (c.pack ? 0u : (1u << 4));
            const uint64_t ad = umma_desc(smem_u32(smem), 32, 128), bd = umma_desc(smem_u32(smem) + 4096, 128, (uint32_t)c.sbo);
            const uint32_t alo = (uint32_t)ad, ahi = (uint32_t)(ad >> 32), blo = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
            uint32_t dcol = 0;
            for (int it = 0; it < c.iters; it++) {
                const bool e = elect_one();
                for (int k = 0; k < c.n_k; k++)
                    if (e) umma_f16_lohi(tmem + dcol, alo, ahi, blo, bhi, idesc, k > 0);
                if (c.commit && e) { for (int q = 0; q < c.commit; q++) umma_commit(bar + 8 + 8 * q); }
                dcol += c.n; if (dcol + c.n > 512u) dcol = 0;
            }
            if (elect_one()) umma_commit(bar);
            __syncwarp();
            mbar_wait(bar, 0, nullptr);
            t1 = clock64();
        }
    } else if (warp <= (uint32_t)c.epi_warps) {
        const uint32_t ew = warp - 1, q = warp & 3;           // TMEM lane quarter = warp id % 4
        const uint32_t sub = ew >> 2, nsub = (c.epi_warps + 3) / 4;
        const uint32_t colsPerLd = (c.shape == 16 ? 16u : 32u) * (c.pack ? 2u : 1u);
        uint32_t acc = 0, v[32];
        for (int i = 0; i < 32; i++) v[i] = 0;
        uint32_t col = sub * colsPerLd;
        for (int it = 0; it < c.iters; it++) {
            for (int l = 0; l < c.lds_per_wait; l++) {
                const uint32_t ta = tmem + ((q * 32u) << 16) + (col & 511u);
                if (c.shape == 16) tmem_ld16(ta, v);
                else if (c.pack) tmem_ld32_pack16(ta, v); else tmem_ld32(ta, v);
                col += nsub * colsPerLd;
            }
            tmem_ld_wait();
            acc ^= v[0] ^ v[7] ^ v[15] ^ v[31];
        }
        t1 = clock64();
        if (acc == 0x12345678u) sink[0] = acc;
    }
    if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main()
{
    unsigned long long* d_out; uint32_t* d_sink;
    cudaMalloc(&d_out, 32 * 8); cudaMalloc(&d_sink, 4);
    cudaFuncSetAttribute(tmem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    std::vector<Cfg> cfgs;
    for (int n : {256, 128})
        for (int nk : {2, 3, 4})
            for (int sbo : {256, 2 * nk * 128})
                for (int commit : {0, 2})
                    for (int fill : {0, 1}) cfgs.push_back({0, 1, 1, 1, nk, 4000, 32, n, sbo, commit, fill});
    printf("epi_warps pack mma lds/wait shape | ld-warp cycles/iter  TMEM-columns*lanes*4B per clk per SM | mma cycles per MMA\n");
    for (const Cfg& c : cfgs) {
        cudaMemset(d_out, 0, 32 * 8);
        for (int rep = 0; rep < 2; rep++) tmem_bw_kernel<<<148, 32 * 17, 200 * 1024>>>(c, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[32]; cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int w = 1; w <= c.epi_warps; w++) if (h[w] > mx) mx = h[w];
        const double cyc = (double)mx / c.iters;
        const double cols = (c.shape == 16 ? 16.0 : 32.0) * (c.pack ? 2 : 1) * c.lds_per_wait * c.epi_warps;   // 32-bit TMEM columns x 32 lanes per iteration
        printf("%9d %4d %3d %8d %5d N=%3d nk=%d sbo=%4d commit=%d fill=%d | %10.1f %10.1f | %8.1f\n", c.epi_warps, c.pack, c.mma, c.lds_per_wait, c.shape, c.n, c.n_k, c.sbo, c.commit, c.fill, cyc,
               c.epi_warps ? cols * 32 * 4 / cyc : 0.0, c.mma ? (double)h[0] / (c.iters * c.n_k) : 0.0);
    }
    return 0;
}
