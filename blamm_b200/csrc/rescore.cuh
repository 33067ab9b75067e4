// Exact rescoring of the candidates that the tensor-core filter let through: the score is re-summed in FP32 in position order
// from the FP32 weights (bit-identical to the in-order sum of the reference's naive path, motif.cpp:225-239; its BLAS path
// re-associates the sums of longer motifs), compared with the exact threshold, checked against the fragment table and the payload
// limit, and appended to the hit list.  Rare work (about 1.7e-4 of all scores).
#pragma once
#include "common.cuh"
#include "filter_tc.cuh"

namespace b200 {

// Raw entries of the tensor-core filter {window, column of chunk 0, mask, mask, column of chunk 1, mask, mask, -}: sign words are
// decoded to masks whose zero bits are the candidates (filter_tc.cuh: sign_words).
// Inverse of filter_tc.cuh: sign_words().  X = sum_b 255 * M_b * 256^b (mod 2^32)  ->  M_0 | M_1 << 8 | M_2 << 16 | M_3 << 24
// (bit 8b + t set <=> accumulator (t, b) negative).  255 M = 256 M - M, so X = -M_0 + (M_0 - M_1) 256 + (M_1 - M_2) 256^2 + ...
// and the bytes peel off from the bottom.  FP32 accumulators only fill b < 2 (the upper bytes repeat them).
__device__ __forceinline__ uint32_t decode_sign_word(bool acc16, uint32_t x)
{
    const uint32_t m0 = (0u - x) & 255u;
    const uint32_t y1 = (x + m0) >> 8;                               // (M_0 - M_1) + (M_1 - M_2) 256 + (M_2 - M_3) 256^2  mod 2^24
    const uint32_t m1 = (m0 - y1) & 255u;
    if (!acc16) return m0 | (m1 << 8) | 0xFFFF0000u;
    const uint32_t y2 = ((y1 - (m0 - m1)) & 0xFFFFFFu) >> 8;         // (M_1 - M_2) + (M_2 - M_3) 256  mod 2^16
    const uint32_t m2 = (m1 - y2) & 255u;
    const uint32_t y3 = ((y2 - (m1 - m2)) & 0xFFFFu) >> 8;           // (M_2 - M_3)  mod 2^8
    const uint32_t m3 = (m2 - y3) & 255u;
    return m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
}

// ---------------------------------------------------------------------------------------------------------
// Fused expand + rescore, column tile by column tile (round 2).
//
// Round 1 expanded the raw entries into a (position, column) candidate list and rescored that with one thread per candidate; ncu
// showed the rescorer L1/TEX bound (l1tex throughput 90 %, 16 sectors per candidate): the lanes of a warp hold candidates of
// different columns, so every weight load of a warp touches up to 32 sectors of the 360 KB FP32 table.  Here the raw blocks
// carry the TAG of the column tile that produced them (filter_tc.cuh: an epilogue warp closes its block when its CTA moves to
// another tile); a CTA takes a batch of consecutive blocks -- in allocation order they are almost always of one tile, because all
// CTAs of the filter work through the items tile by tile -- loads THAT tile's FP32 weights and column records into shared memory
// once, and scores the batch's raw entries straight from them: a warp per block, two entries per lane in flight (prefetching the
// next block's as well spills at the 64 registers two CTAs per SM allow: 1.6 instead of 1.07 ms), the window's codes loaded once per
// entry, every weight read an LDS.  Every lane scores its entry's first candidate in
// the same round (96 % of the entries have one); further candidates go through a per-warp queue and are scored 32 at a time.
// The candidate list and its two kernels are gone.  Scores are the same in-order FP32 sums, hits leave through warp staging.
// Tiles whose weights exceed the shared-memory budget (256 columns of > 36 positions) are scored from global memory.
// Measured (100 Mbp x 1800 columns, 3.0e7 candidates): 1.07 ms against 0.33 + 1.00 ms for expand + rescore; with half of the
// sequence soft-masked (4.1e8 candidates) 12.4 against 12.9 ms.
// ---------------------------------------------------------------------------------------------------------
// FP32 score of one window against one column: the weights of the window's letters added strictly in position order (16 at a time:
// all loads first, then the additions); masked positions add nothing (the reference's BLAS-path semantics of lower case)
template <bool MASKED>
__device__ __forceinline__ float score_in_order(const float* wp, uint32_t L, const uint32_t (&codes)[4], const uint32_t (&zm)[2])
{
    float s = 0.0f;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if ((uint32_t)(16 * g) < L) {
            const uint32_t rr = codes[g], n = min(16u, L - 16u * g);
            float wv[16];
#pragma unroll
            for (uint32_t t = 0; t < 16; t++) wv[t] = (t < n) ? wp[4 * (16 * g + t) + ((rr >> (2 * t)) & 3u)] : 0.0f;
#pragma unroll
            for (uint32_t t = 0; t < 16; t++)
                if (t < n && !(MASKED && ((zm[g >> 1] >> (16 * (g & 1) + t)) & 1u))) s += wv[t];
        }
    }
    return s;
}

constexpr uint32_t kFuseThreads = 512;
constexpr uint32_t kFuseBatch   = 256;                 // raw blocks per work item (16,384 entry slots)
constexpr uint32_t kFuseRounds  = 2;                   // warp rounds of hits staged per global atomic
constexpr uint32_t kFuseMaxW    = 9728;                // positions of weights a tile may hold in shared memory (152 KB)
__host__ __device__ constexpr size_t fuse_smem_bytes(uint32_t max_w) { return (size_t)max_w * 16 + 256 * 16 + (size_t)(kFuseThreads / 32) * 32 * kFuseRounds * 16; }
struct ColRec { uint32_t woff; uint32_t len; float thr; uint32_t orig; };

template <bool MASKED>
__global__ void __launch_bounds__(kFuseThreads)
rescore_tile_kernel(MotifDev md, BlockDev blk, const uint32_t* __restrict__ raw, const uint32_t* __restrict__ blk_count,
                    const uint32_t* __restrict__ blk_tag, const unsigned int* __restrict__ n_blocks_ptr, uint32_t blk_cap,
                    unsigned int* work_counter, unsigned long long* n_cand, uint32_t max_w, HitSink sink)
{
    if ((__ldg(blk.has_zero) != 0) != MASKED) return;
    extern __shared__ __align__(16) uint8_t fuse_smem[];
    float4* s_w = reinterpret_cast<float4*>(fuse_smem);                                       // weights of the tile: max_w positions
    ColRec* s_col = reinterpret_cast<ColRec*>(fuse_smem + (size_t)max_w * 16);                 // its <= 256 column records
    constexpr uint32_t kRounds = kFuseRounds;
    b200scan_hit* s_hits = reinterpret_cast<b200scan_hit*>(s_col + 256);                      // [warp][32 * kRounds]
    __shared__ uint32_t s_item, s_next_tag;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    b200scan_hit* st = s_hits + wib * 32 * kRounds;
    __shared__ uint2 s_queue[kFuseThreads / 32][64];               // per warp: (window, relative column) of the candidates beyond an entry's first
    uint2* qbuf = s_queue[wib];
    uint32_t qn = 0;
    uint32_t n_st = 0, round = 0;
    unsigned long long my_cand = 0;
    auto flush = [&]() {
        __syncwarp();
        unsigned long long o = 0;
        if (lane == 0) o = atomicAdd(sink.n_hits, (unsigned long long)n_st);
        o = __shfl_sync(0xffffffffu, o, 0);
        for (uint32_t k0 = 0; k0 < n_st; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool stored = k < n_st && o + k < sink.cap;
            if (stored) store_hit(sink, o + k, st[k]);
            if (sink.bucket_cnt) {
                const uint32_t sm = __ballot_sync(0xffffffffu, stored);
                if (stored) count_hit_bucket(sink, sm, (uint32_t)st[k].pos);
            }
        }
        __syncwarp();
        n_st = 0; round = 0;
    };
    const uint32_t nb = min(*n_blocks_ptr, blk_cap);
    const uint4* ent = reinterpret_cast<const uint4*>(raw);
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t b0 = s_item * kFuseBatch;
        if (b0 >= nb) break;
        const uint32_t b1 = min(nb, b0 + kFuseBatch);
        // the batch's tiles, one after the other in increasing tag order (almost always one tile; two or three where the filter's
        // CTAs changed tile) -- every tile exactly once, however many there are
        uint32_t last_tag = 0;                                // tags are >= 1 (n_cols >= 1)
        for (;;) {
            if (threadIdx.x == 0) s_next_tag = 0xffffffffu;
            __syncthreads();
            for (uint32_t b = b0 + threadIdx.x; b < b1; b += kFuseThreads) {
                if (__ldg(blk_count + b) == 0) continue;      // spare / unused reservations carry no tag
                const uint32_t t = __ldg(blk_tag + b);
                if (t > last_tag) atomicMin(&s_next_tag, t);
            }
            __syncthreads();
            const uint32_t tag = s_next_tag;
            if (tag == 0xffffffffu) break;
            const uint32_t col0 = tag >> 9, ncol = tag & 511u;
            const uint32_t wbase = __ldg(md.woff + col0);
            const uint32_t nW = __ldg(md.woff + col0 + ncol - 1) + __ldg(md.len + col0 + ncol - 1) - wbase;
            const bool in_smem = nW <= max_w;
            __syncthreads();                                  // (everybody has read s_next_tag and is done with the previous tile's tables)
            if (in_smem) for (uint32_t i = threadIdx.x; i < nW; i += kFuseThreads) s_w[i] = __ldg(md.w + wbase + i);
            for (uint32_t c = threadIdx.x; c < ncol; c += kFuseThreads) {
                ColRec r; r.woff = __ldg(md.woff + col0 + c) - wbase; r.len = __ldg(md.len + col0 + c); r.thr = __ldg(md.thr + col0 + c); r.orig = __ldg(md.orig + col0 + c);
                s_col[c] = r;
            }
            __syncthreads();
            const float* wsm = reinterpret_cast<const float*>(s_w);
            const float* wgl = reinterpret_cast<const float*>(md.w + wbase);
            // one round: every lane scores column c (relative to the tile; 0xffffffff = nothing) of the window at pos; hits are staged
            auto score_round = [&](uint32_t c, uint32_t pos, const uint32_t (&codes)[4], const uint32_t (&zm)[2]) {
                bool hit = false; uint32_t colo = 0; float sc = 0.0f;
                if (c < ncol) {                                      // (a padding column can never be a candidate; belt and braces)
                    const ColRec r = s_col[c];
                    sc = in_smem ? score_in_order<MASKED>(wsm + 4 * r.woff, r.len, codes, zm) : score_in_order<MASKED>(wgl + 4 * r.woff, r.len, codes, zm);
                    hit = (pos < blk.n_payload) && !(sc < r.thr);
                    if (hit) hit = window_in_fragment(blk, pos, r.len);
                    colo = r.orig;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    b200scan_hit hh; hh.pos = pos; hh.col = colo; hh.score = sc;
                    st[n_st + __popc(m & ((1u << lane) - 1u))] = hh;
                }
                n_st += __popc(m);
                if (++round == kRounds) flush();
            };
            // n queued candidates starting at qbuf[from]: one per lane, with their window's codes loaded again (L1 / L2 hits)
            auto drain_queue = [&](uint32_t from, uint32_t n) {
                uint32_t c = 0xffffffffu, pos = 0, codes[4] = {0u, 0u, 0u, 0u}, zm[2] = {0u, 0u};
                if (lane < n) {
                    const uint2 e = qbuf[from + lane];
                    pos = e.x; c = e.y;
                    load_window_codes(blk.codes, pos, codes);
                    if (MASKED) load_window_zmask(blk.zmask, pos, zm);
                }
                __syncwarp();
                score_round(c, pos, codes, zm);
            };
            // one warp per raw block of this tile and pass; a lane takes the block's entries `lane` and `lane + 32`: both are loaded up
            // front, then both windows' codes (the kernel is bound by memory latency: bytes in flight per thread are what count)
            for (uint32_t b = b0 + wib; b < b1; b += kFuseThreads / 32) {
                if (__ldg(blk_tag + b) != tag) continue;                                        // (warp-uniform)
                const uint32_t cnt = __ldg(blk_count + b);
                if (cnt == 0) continue;
                uint4 x[2], y[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    x[h] = make_uint4(0u, 0u, kAllNegative, kAllNegative); y[h] = make_uint4(0u, kAllNegative, kAllNegative, 0u);
                    const uint32_t e = lane + 32 * h;
                    if (e < cnt) { x[h] = __ldg(ent + 2 * ((size_t)b * kRawBlock + e)); y[h] = __ldg(ent + 2 * ((size_t)b * kRawBlock + e) + 1); }
                }
                uint32_t zz[2][4], codes2[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}}, zm2[2][2] = {{0u, 0u}, {0u, 0u}};
                bool acc16h[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    acc16h[h] = !(x[h].y & kRawFp32Flag);
                    zz[h][0] = ~decode_sign_word(acc16h[h], x[h].z); zz[h][1] = ~decode_sign_word(acc16h[h], x[h].w);
                    zz[h][2] = ~decode_sign_word(acc16h[h], y[h].y); zz[h][3] = ~decode_sign_word(acc16h[h], y[h].z);
                    if ((zz[h][0] | zz[h][1] | zz[h][2] | zz[h][3]) != 0u) {
                        load_window_codes(blk.codes, x[h].x, codes2[h]);
                        if (MASKED) load_window_zmask(blk.zmask, x[h].x, zm2[h]);
                    }
                    my_cand += __popc(zz[h][0]) + __popc(zz[h][1]) + __popc(zz[h][2]) + __popc(zz[h][3]);
                }
                // Every lane scores the FIRST candidate of its entry at once (an entry has one candidate in 96 % of the cases); further
                // candidates go to the warp's queue and are scored 32 at a time -- a second round for the whole warp because one or two
                // lanes have a second candidate would double the work of the usual case.
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (h == 1 && cnt <= 32) break;
                    const bool acc16 = acc16h[h];
                    uint32_t (&z)[4] = zz[h];
                    const uint32_t first[2] = {(x[h].y & ~kRawFp32Flag) - col0, (y[h].x & ~kRawFp32Flag) - col0};        // columns relative to the tile
                    const uint32_t pos = x[h].x;
                    uint32_t q = 0;
                    auto next_col = [&]() -> uint32_t {              // pops the lane's next candidate column (0xffffffff: none)
                        while (q < 4 && z[q] == 0u) q++;
                        if (q >= 4) return 0xffffffffu;
                        const uint32_t bit = __ffs(z[q]) - 1; z[q] &= z[q] - 1;
                        const uint32_t w = q & 1u;
                        return acc16 ? first[q >> 1] + 32 * w + 4 * (bit & 7u) + (bit >> 3) : first[q >> 1] + 16 * w + 2 * (bit & 7u) + (bit >> 3);
                    };
                    score_round(next_col(), pos, codes2[h], zm2[h]);
                    while (__any_sync(0xffffffffu, (z[0] | z[1] | z[2] | z[3]) != 0u)) {
                        const uint32_t c = next_col();
                        const unsigned m = __ballot_sync(0xffffffffu, c != 0xffffffffu);
                        if (c != 0xffffffffu) qbuf[qn + __popc(m & ((1u << lane) - 1u))] = make_uint2(pos, c);
                        qn += __popc(m);
                        __syncwarp();
                        if (qn >= 32) { qn -= 32; drain_queue(qn, 32); }
                    }
                }
            }
            if (qn) { drain_queue(0, qn); qn = 0; }                  // the queued candidates belong to this tile's tables
            last_tag = tag;
        }
        __syncthreads();
    }
    if (n_st) flush();
    // candidates seen (b200scan_timing.n_candidates): one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_cand += __shfl_xor_sync(0xffffffffu, my_cand, d);
    if (lane == 0 && my_cand) atomicAdd(n_cand, my_cand);
}

} // namespace b200
