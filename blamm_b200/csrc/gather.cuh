// Gather-add scoring kernel: score(g, col) = sum_{j < L} W_col[j][code[g+j]], added in position order in FP32,
// i.e. the naive in-order sum: bit-identical to the reference's naive path (Motif::getScore, motif.cpp:225-239) and to its sgemm
// wherever the BLAS keeps one in-order FMA chain per output (the zero products add exactly -- SURVEY.md 7.3 [probe]); OpenBLAS
// re-associates the sums of motifs longer than ~18 positions, and there the reference's BLAS path differs from its own naive path by
// <= 3.8e-6.  Replaces, for one block,
//   the w-iteration loop { Matrix::sgemm_batch ; extractOccurrences }           pwmscan.cpp:259-264, :104-133
// without ever materialising the one-hot matrix S or the score matrix R.
//
// Roofline: shared-memory lookups.  One LDS per (window, column, position); 32 lanes/clk/SM.
// Layout: a CTA owns a tile of columns whose FP32 weights (float4 per position) sit in shared memory and
// sweeps a span of windows; each thread keeps its window's 64 characters of 2-bit codes in 4 registers.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kGatherThreads = 256;
constexpr int kGatherSpan    = 4096;        // windows per CTA (16 per thread, one at a time)

struct GatherTile {             // host-built, one per column tile
    uint32_t col0, n_cols;      // sorted column range
    uint32_t w0, n_w;           // range in MotifDev::w (float4 units)
};

template <bool ZMASK>
__global__ void __launch_bounds__(kGatherThreads)
gather_scan_kernel(MotifDev md, BlockDev blk, const GatherTile* __restrict__ tiles, HitSink sink,
                   int run_if_zero /* 0: run only if !has_zero, 1: only if has_zero, 2: always */)
{
    extern __shared__ float4 smem_w[];      // n_w float4, then per-column meta
    const uint32_t hz = __ldg(blk.has_zero);
    if (run_if_zero == 0 && hz != 0) return;
    if (run_if_zero == 1 && hz == 0) return;

    const GatherTile tile = tiles[blockIdx.y];
    uint32_t* s_off = reinterpret_cast<uint32_t*>(smem_w + tile.n_w);   // local float4 offset per column
    uint32_t* s_len = s_off + tile.n_cols;
    float*    s_thr = reinterpret_cast<float*>(s_len + tile.n_cols);

    for (uint32_t i = threadIdx.x; i < tile.n_w; i += blockDim.x) smem_w[i] = __ldg(md.w + tile.w0 + i);
    for (uint32_t i = threadIdx.x; i < tile.n_cols; i += blockDim.x) {
        s_off[i] = __ldg(md.woff + tile.col0 + i) - tile.w0;
        s_len[i] = __ldg(md.len + tile.col0 + i);
        s_thr[i] = __ldg(md.thr + tile.col0 + i);
    }
    __syncthreads();

    const uint32_t span0 = blockIdx.x * kGatherSpan;
    for (uint32_t it = 0; it < kGatherSpan / kGatherThreads; it++) {
        const uint32_t pos = span0 + it * kGatherThreads + threadIdx.x;
        if (span0 + it * kGatherThreads >= blk.n_payload) break;     // block-uniform
        const bool live = pos < blk.n_payload;
        uint32_t codes[4], zm[2] = {0, 0};
        load_window_codes(blk.codes, live ? pos : 0, codes);
        if (ZMASK) load_window_zmask(blk.zmask, live ? pos : 0, zm);

        for (uint32_t c = 0; c < tile.n_cols; c++) {
            const uint32_t L = s_len[c];
            const float* wp = reinterpret_cast<const float*>(smem_w + s_off[c]);
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((uint32_t)(16 * q) < L) {
                    uint32_t r = codes[q];
                    uint32_t z = ZMASK ? ((zm[q >> 1] >> (16 * (q & 1))) & 0xFFFFu) : 0u;
                    const uint32_t n = min(16u, L - 16u * q);
                    const float* wq = wp + 64 * q;
#pragma unroll 4
                    for (uint32_t t = 0; t < n; t++) {
                        float w = wq[4 * t + (r & 3u)];
                        if (ZMASK) { if (z & 1u) w = 0.0f; z >>= 1; }
                        s += w;                       // in-order FP32 add (the reference's naive path; its sgemm up to re-association)
                        r >>= 2;
                    }
                }
            }
            bool hit = live && !(s < s_thr[c]);        // `if (thisScore < threshold) continue;` pwmscan.cpp:116
            if (__any_sync(0xffffffffu, hit)) {
                if (hit) hit = window_in_fragment(blk, pos, L);
                emit_hits_warp(hit, pos, __ldg(md.orig + tile.col0 + c), s, sink);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Empirical score histograms (`blamm hist -e`): the same exact scores, with the reference's histogram epilogue
//   Histogram::extractObsScore + ScoreHistogram::addObservation  (hist.cpp:70-93, motif.h:96-102):
//   every window that lies inside one fragment adds 1 to bin  clamp(int((score - min) / width), 0, bins-1)  of its column.
// A CTA keeps the histograms of its column tile in shared memory (u32, one smem atomic per score) over a long span
// of windows and flushes the non-empty bins with 64-bit global atomics once.
// ---------------------------------------------------------------------------------------------------------
constexpr int kHistSpan = 32768;            // windows per CTA

template <bool ZMASK>
__global__ void __launch_bounds__(kGatherThreads)
gather_hist_kernel(MotifDev md, BlockDev blk, const GatherTile* __restrict__ tiles, const float* __restrict__ hmin,
                   const float* __restrict__ hwidth, uint32_t num_bins, unsigned long long* __restrict__ hist, int run_if_zero)
{
    extern __shared__ float4 smem_w[];      // n_w float4 | per-column meta | n_cols * num_bins u32 counters
    const uint32_t hz = __ldg(blk.has_zero);
    if (run_if_zero == 0 && hz != 0) return;
    if (run_if_zero == 1 && hz == 0) return;

    const GatherTile tile = tiles[blockIdx.y];
    uint32_t* s_off = reinterpret_cast<uint32_t*>(smem_w + tile.n_w);
    uint32_t* s_len = s_off + tile.n_cols;
    float*    s_min = reinterpret_cast<float*>(s_len + tile.n_cols);
    float*    s_wid = s_min + tile.n_cols;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_wid + tile.n_cols);

    for (uint32_t i = threadIdx.x; i < tile.n_w; i += blockDim.x) smem_w[i] = __ldg(md.w + tile.w0 + i);
    for (uint32_t i = threadIdx.x; i < tile.n_cols; i += blockDim.x) {
        s_off[i] = __ldg(md.woff + tile.col0 + i) - tile.w0;
        s_len[i] = __ldg(md.len + tile.col0 + i);
        s_min[i] = __ldg(hmin + tile.col0 + i);
        s_wid[i] = __ldg(hwidth + tile.col0 + i);
    }
    for (uint32_t i = threadIdx.x; i < tile.n_cols * num_bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();

    const uint32_t span0 = blockIdx.x * kHistSpan;
    for (uint32_t it = 0; it < kHistSpan / kGatherThreads; it++) {
        const uint32_t pos = span0 + it * kGatherThreads + threadIdx.x;
        if (span0 + it * kGatherThreads >= blk.n_payload) break;     // block-uniform
        if (pos >= blk.n_payload) continue;
        uint32_t codes[4], zm[2] = {0, 0};
        load_window_codes(blk.codes, pos, codes);
        if (ZMASK) load_window_zmask(blk.zmask, pos, zm);
        const uint32_t remaining = fragment_remaining(blk, pos);
        for (uint32_t c = 0; c < tile.n_cols; c++) {
            const uint32_t L = s_len[c];
            if (L > remaining) continue;                              // hist.cpp:87-88
            const float* wp = reinterpret_cast<const float*>(smem_w + s_off[c]);
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((uint32_t)(16 * q) < L) {
                    uint32_t r = codes[q];
                    uint32_t z = ZMASK ? ((zm[q >> 1] >> (16 * (q & 1))) & 0xFFFFu) : 0u;
                    const uint32_t n = min(16u, L - 16u * q);
                    const float* wq = wp + 64 * q;
#pragma unroll 4
                    for (uint32_t t = 0; t < n; t++) {
                        float w = wq[4 * t + (r & 3u)];
                        if (ZMASK) { if (z & 1u) w = 0.0f; z >>= 1; }
                        s += w;
                        r >>= 2;
                    }
                }
            }
            int bin = (int)__fdiv_rn(s - s_min[c], s_wid[c]);          // int((score - minScore) / width), motif.h:97
            bin = max(0, bin);
            bin = min((int)num_bins - 1, bin);
            atomicAdd(&s_hist[c * num_bins + bin], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile.n_cols * num_bins; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(hist + (size_t)__ldg(md.orig + tile.col0 + i / num_bins) * num_bins + (i % num_bins), (unsigned long long)v);
    }
}

} // namespace b200
