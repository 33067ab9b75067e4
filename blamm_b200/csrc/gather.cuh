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

// One more observation in bin `bin` of a shared-memory histogram row.  The 32 lanes of a warp score the SAME column for 32
// neighbouring windows, and the scores of a column crowd into a few dozen bins: a plain atomicAdd serialises on the lanes
// that share a bin.  With AGG the lanes that share a bin elect one of them to add their count (every lane of the warp must
// call this; `ok` = this lane has an observation).
template <bool AGG>
__device__ __forceinline__ void hist_add(uint32_t* row, int bin, bool ok)
{
    if (AGG) {
        const uint32_t peers = __match_any_sync(0xffffffffu, ok ? bin : -1);
        if (ok && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(row + bin, (uint32_t)__popc(peers));
    } else if (ok) atomicAdd(row + bin, 1u);
}

template <bool ZMASK, bool AGG = false>
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
        const bool live = pos < blk.n_payload;
        if (!AGG && !live) continue;                                  // (AGG: the whole warp stays together for the vote in hist_add)
        uint32_t codes[4], zm[2] = {0, 0};
        load_window_codes(blk.codes, live ? pos : 0, codes);
        if (ZMASK) load_window_zmask(blk.zmask, live ? pos : 0, zm);
        const uint32_t remaining = live ? fragment_remaining(blk, pos) : 0u;
        for (uint32_t c = 0; c < tile.n_cols; c++) {
            const uint32_t L = s_len[c];
            const bool ok = L <= remaining;                           // hist.cpp:87-88
            if (!AGG && !ok) continue;
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
            hist_add<AGG>(s_hist + c * num_bins, bin, ok);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile.n_cols * num_bins; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(hist + (size_t)__ldg(md.orig + tile.col0 + i / num_bins) * num_bins + (i % num_bins), (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same histograms with the scoring loop turned inside out (B200SCAN_HIST_KERNEL=3, the default; 2 = without the rotation
// of the bin counts, ROT = false).
//   * Weights of a column tile lie in shared memory as rows of POSITIONS: W[t][column] (float4 = A,C,G,T), zero where a column
//     is shorter than the row index, plus one all-zero row.  For a window the byte offset of "row t, letter code[t]" is computed
//     ONCE (P[t] = t * rowBytes + 4 * code[t]; a masked position points at the zero row) and serves every column of the tile:
//     the inner loop is one LDS with an immediate column offset and one FADD per (position, column), eight columns -- eight
//     independent in-order sums -- at a time.  Adding +0.0f behind a column's last position leaves its sum bit for bit as it is
//     (an in-order sum that starts at +0.0f is never -0.0f).
//   * ROT: the lanes of a warp count different columns in the same instruction (see the epilogue); without it hist_add<true>.
// Roofline: shared-memory lookups, 32 lanes/clk/SM (one LDS per lane, position and column).
// ---------------------------------------------------------------------------------------------------------
constexpr int kHist2Threads = 256;
constexpr int kHist2U       = 8;            // columns scored together (independent FADD chains per thread)

struct HistTile2 {              // host-built, one per column tile (columns sorted by length)
    uint32_t col0, n_cols;      // sorted column range
    uint32_t n_pad;             // n_cols rounded up to kHist2U (padding columns are never counted)
    uint32_t max_len;           // length of the tile's longest column = position rows of W
};
__host__ __device__ inline size_t hist2_smem_bytes(uint32_t n_cols, uint32_t max_len, uint32_t num_bins)
{
    const size_t n_pad = (n_cols + kHist2U - 1) / kHist2U * kHist2U;
    return (size_t)(max_len + 1) * n_pad * 16 + n_pad * 12 + (size_t)n_cols * num_bins * 4;
}

template <bool ZMASK, bool ROT>
__global__ void __launch_bounds__(kHist2Threads, 2)
gather_hist2_kernel(MotifDev md, BlockDev blk, const HistTile2* __restrict__ tiles, const float* __restrict__ hmin,
                    const float* __restrict__ hwidth, uint32_t num_bins, unsigned long long* __restrict__ hist, int run_if_zero)
{
    extern __shared__ float4 smem_w[];      // W[max_len + 1][n_pad] | len[n_pad] | min[n_pad] | width[n_pad] | n_cols * num_bins u32 counters
    const uint32_t hz = __ldg(blk.has_zero);
    if (run_if_zero == 0 && hz != 0) return;
    if (run_if_zero == 1 && hz == 0) return;

    const HistTile2 tile = tiles[blockIdx.y];
    const uint32_t np = tile.n_pad, ML = tile.max_len;
    uint32_t* s_len  = reinterpret_cast<uint32_t*>(smem_w + (size_t)(ML + 1) * np);
    float*    s_min  = reinterpret_cast<float*>(s_len + np);
    float*    s_wid  = s_min + np;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_wid + np);

    for (uint32_t i = threadIdx.x; i < (ML + 1) * np; i += blockDim.x) {
        const uint32_t t = i / np, c = i - t * np;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (c < tile.n_cols && t < __ldg(md.len + tile.col0 + c)) v = __ldg(md.w + __ldg(md.woff + tile.col0 + c) + t);
        smem_w[i] = v;
    }
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
        const bool real = i < tile.n_cols;
        s_len[i] = real ? __ldg(md.len + tile.col0 + i) : 0xFFFFFFFFu;      // a padding column fits no fragment
        s_min[i] = real ? __ldg(hmin + tile.col0 + i) : 0.0f;
        s_wid[i] = real ? __ldg(hwidth + tile.col0 + i) : 1.0f;
    }
    for (uint32_t i = threadIdx.x; i < tile.n_cols * num_bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();

    const uint32_t rowBytes = np * 16u;
    const char* const wbase = reinterpret_cast<const char*>(smem_w);
    const uint32_t span0 = blockIdx.x * kHistSpan;
    for (uint32_t it = 0; it < kHistSpan / kHist2Threads; it++) {
        const uint32_t pos = span0 + it * kHist2Threads + threadIdx.x;
        if (span0 + it * kHist2Threads >= blk.n_payload) break;      // block-uniform
        const bool live = pos < blk.n_payload;                        // (dead lanes stay with their warp for the vote in hist_add)
        uint32_t codes[4], zm[2] = {0, 0};
        load_window_codes(blk.codes, live ? pos : 0, codes);
        if (ZMASK) load_window_zmask(blk.zmask, live ? pos : 0, zm);
        const uint32_t remaining = live ? fragment_remaining(blk, pos) : 0u;
        uint32_t P[kMaxLen];                                          // byte offset of (row t, this window's letter) in W
#pragma unroll
        for (int t = 0; t < kMaxLen; t++) {
            const uint32_t code = (codes[t >> 4] >> (2 * (t & 15))) & 3u;
            P[t] = (uint32_t)t * rowBytes + 4u * code;
            if (ZMASK && ((zm[t >> 5] >> (t & 31)) & 1u)) P[t] = ML * rowBytes;      // contributes exactly 0
        }
        for (uint32_t c0 = 0; c0 < np; c0 += kHist2U) {
            float s[kHist2U];
#pragma unroll
            for (int u = 0; u < kHist2U; u++) s[u] = 0.0f;
            const char* const wc = wbase + c0 * 16u;
#pragma unroll
            for (int t = 0; t < kMaxLen; t++) {
                if ((uint32_t)t >= ML) break;                         // warp-uniform
                const char* const a = wc + P[t];
#pragma unroll
                for (int u = 0; u < kHist2U; u++) s[u] += *reinterpret_cast<const float*>(a + 16 * u);      // in-order FP32 add
            }
            if (ROT) {
                // The lanes of a warp hold the same eight columns, and a column's scores crowd into a few dozen bins: adding column u in
                // step u from all 32 lanes serialises the shared-memory atomics on the lanes that share a bin.  So lane l takes column
                // (k + l) mod 8 in step k -- eight different histogram rows per instruction -- after rotating its sums by l with three
                // rounds of selects (a barrel shifter in registers; the three predicates are per-thread constants).
                static_assert(kHist2U == 8, "the rotation below is written for eight columns");
                const uint32_t lane = threadIdx.x & 31u;
                float r1[8], r2[8], r4[8];
#pragma unroll
                for (int k = 0; k < 8; k++) r1[k] = (lane & 1u) ? s[(k + 1) & 7] : s[k];
#pragma unroll
                for (int k = 0; k < 8; k++) r2[k] = (lane & 2u) ? r1[(k + 2) & 7] : r1[k];
#pragma unroll
                for (int k = 0; k < 8; k++) r4[k] = (lane & 4u) ? r2[(k + 4) & 7] : r2[k];      // r4[k] = s[(k + lane) & 7]
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t c = c0 + ((lane + (uint32_t)k) & 7u);
                    const bool ok = s_len[c] <= remaining;            // hist.cpp:87-88 (false for padding columns and dead lanes)
                    int bin = (int)__fdiv_rn(r4[k] - s_min[c], s_wid[c]);      // int((score - minScore) / width), motif.h:97
                    bin = max(0, bin);
                    bin = min((int)num_bins - 1, bin);
                    if (ok) atomicAdd(s_hist + c * num_bins + bin, 1u);
                }
            } else {
#pragma unroll
                for (int u = 0; u < kHist2U; u++) {
                    const uint32_t c = c0 + u;
                    const bool ok = s_len[c] <= remaining;            // hist.cpp:87-88 (false for padding columns and dead lanes)
                    int bin = (int)__fdiv_rn(s[u] - s_min[c], s_wid[c]);       // int((score - minScore) / width), motif.h:97
                    bin = max(0, bin);
                    bin = min((int)num_bins - 1, bin);
                    hist_add<true>(s_hist + c * num_bins, bin, ok);
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile.n_cols * num_bins; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(hist + (size_t)__ldg(md.orig + tile.col0 + i / num_bins) * num_bins + (i % num_bins), (unsigned long long)v);
    }
}

} // namespace b200
