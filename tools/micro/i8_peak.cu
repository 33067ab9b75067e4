// Microbenchmark: dense tcgen05.mma peak of the tensor pipe the filter kernel runs on, measured the way SURVEY.md 8(d) asks for
// ("peak(kind) measured on the same box by a plain GEMM of that MMA kind"): every SM runs a K loop of M = 128, N = 256 MMAs out of
// shared memory into two alternating TMEM accumulators, nothing else (no epilogue, no global traffic).  Two kinds: kind::i8
// (K = 32 per instruction, S32 accumulators -- the filter's default) and kind::f16 (K = 16, FP32 accumulators) as a cross-check
// against the cuBLAS bf16 figure of MEASURED_PEAKS.json.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o i8_peak tools/micro/i8_peak.cu ; run on a B200.
// Output: one JSON line {"i8_tops": .., "f16_tflops": .., "sm_mhz_i8": .., ...} (tools/gpu_r2_call1.sh keeps it under profiles/).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../blamm_b200/csrc/filter_tc.cuh"
using namespace b200;

constexpr int kKSteps = 8;           // K steps per accumulation chain (a chain = one 128 x 256 x (8 * K) tile product)

template <bool I8>
__global__ void __launch_bounds__(128, 1) peak_kernel(int chains, unsigned long long* cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) unsigned long long bar_s[2];
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    // operands: small non-zero values so the data path toggles (INT8: bytes 0..3 or -128..-125; FP16: +-1, +-0.5)
    for (uint32_t i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) {
        const uint32_t h = i * 2654435761u;
        ((uint32_t*)smem)[i] = I8 ? (h & 0x83838383u) : ((h & 0x80008000u) | 0x38003C00u);
    }
    const uint32_t bar = smem_u32(&bar_s[0]);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        // K-major no-swizzle operands (filter_tc.cuh: umma_desc): a K step is two 16-byte chunks per row.
        // A: 128 rows -> per K step 128 x 32 B = 4 KB ([chunk][row] blocks: LBO = 2048 B between chunks, SBO = 128 B between 8-row groups)
        // B: 256 rows -> per K step 8 KB (LBO = 4096 B, SBO = 128 B)
        const uint32_t idesc = (I8 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad = umma_desc(smem_u32(smem), 2048, 128), bd = umma_desc(smem_u32(smem) + 32768, 4096, 128);
        const uint32_t alo0 = (uint32_t)ad, ahi = (uint32_t)(ad >> 32), blo0 = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
        const unsigned long long t0 = clock64();
        if (elect_one()) {
#pragma unroll 1
            for (int c = 0; c < chains; c++) {
                const uint32_t d = tmem + (c & 1) * 256;
                uint32_t alo = alo0, blo = blo0;
                umma_x<false, I8>(d, alo, ahi, blo, bhi, idesc, 0u);
#pragma unroll
                for (int k = 1; k < kKSteps; k++) {
                    alo += 4096 >> 4; blo += 8192 >> 4;
                    umma_x<false, I8>(d, alo, ahi, blo, bhi, idesc, 1u);
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0, nullptr);
        const unsigned long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <bool I8> static int run(const char* name, int sms, int chains, double* rate_out, double* mhz_out)
{
    unsigned long long* d_cyc; cudaMalloc(&d_cyc, sms * 8);
    const int smem = 96 * 1024;
    cudaFuncSetAttribute(peak_kernel<I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best_ms = 1e30; unsigned long long cyc_at_best = 0;
    for (int rep = 0; rep < 6; rep++) {                     // rep 0-1 warm up the clocks
        cudaEventRecord(e0);
        peak_kernel<I8><<<sms, 128, smem>>>(chains, d_cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", name, cudaGetErrorString(e)); return 1; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> h(sms); cudaMemcpy(h.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
        if (rep >= 2 && ms < best_ms) { best_ms = ms; cyc_at_best = mx; }
    }
    const double ops = 2.0 * 128 * 256 * (I8 ? 32 : 16) * kKSteps * (double)chains * sms;
    *rate_out = ops / (best_ms * 1e-3) / 1e12;
    *mhz_out = cyc_at_best / (best_ms * 1e-3) / 1e6;         // SM clock seen by the kernel itself (clock64 over the event time)
    cudaFree(d_cyc);
    return 0;
}

int main(int argc, char** argv)
{
    int dev = argc > 2 ? atoi(argv[2]) : 0; cudaSetDevice(dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int sms = p.multiProcessorCount;
    const int chains = argc > 1 ? atoi(argv[1]) : 4000;      // 4000 chains x 8 MMAs x 128 cycles = 4.1 M cycles ~ 2 ms per launch
    double i8 = 0, f16 = 0, mhz8 = 0, mhz16 = 0;
    if (run<true>("i8", sms, chains, &i8, &mhz8)) return 1;
    if (run<false>("f16", sms, chains, &f16, &mhz16)) return 1;
    printf("{\"i8_tops\": %.1f, \"f16_tflops\": %.1f, \"sm_mhz_i8\": %.0f, \"sm_mhz_f16\": %.0f, \"sms\": %d, \"shape\": \"M128 N256 K%d/%d per tcgen05.mma, %d-step chains, %d chains per SM\", "
           "\"how\": \"tools/micro/i8_peak.cu: dense tcgen05.mma loop out of shared memory on every SM, best of 4 launches, CUDA events\"}\n",
           i8, f16, mhz8, mhz16, sms, 32, 16, kKSteps, chains);
    return 0;
}
