// Exact rescoring of the candidates that the tensor-core filter let through.  One thread per candidate:
// the score is re-summed in FP32 in position order from the FP32 weights (bit-identical to the reference's
// sgemm chain, see gather.cuh), compared with the exact threshold, checked against the fragment table and
// the payload limit, and appended to the hit list.  Rare work (about 1e-4 of all scores), L2-resident.
#pragma once
#include "common.cuh"
#include "filter_tc.cuh"

namespace b200 {

// Raw entries of the tensor-core filter -> (position, sorted column) candidates.  One thread per entry slot
// {window, column of chunk 0, mask, mask, column of chunk 1, mask, mask, -}; sign words are decoded to masks whose zero bits are the
// candidates (filter_tc.cuh: sign_words).  A warp's candidates are staged in shared memory and appended with one global atomic per
// >= 256 of them.
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

__global__ void __launch_bounds__(256)
expand_kernel(const uint32_t* __restrict__ raw, const uint32_t* __restrict__ blk_count, const unsigned int* __restrict__ n_blocks_ptr,
              uint32_t blk_cap, Cand* __restrict__ cand, unsigned long long* n_cand, unsigned long long cand_cap,
              const uint32_t* __restrict__ has_zero)
{
    constexpr uint32_t kStage = 640;
    __shared__ Cand s_stage[8][kStage];   // per-warp staging: one global atomic per >= 256 candidates
    (void)has_zero;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t nb = min(*n_blocks_ptr, blk_cap);
    Cand* st = s_stage[wib];
    uint32_t n = 0;
    auto flush = [&]() {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(n_cand, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t s = lane; s < n; s += 32)
            if (base + s < cand_cap) cand[base + s] = st[s];
        __syncwarp();
        n = 0;
    };
    auto column = [](bool acc16, uint32_t first, uint32_t w, uint32_t bit) {
        return acc16 ? first + 32 * w + 4 * (bit & 7u) + (bit >> 3) : first + 16 * w + 2 * (bit & 7u) + (bit >> 3);
    };
    const unsigned long long slots = (unsigned long long)nb * kRawBlock;
    const uint4* ent = reinterpret_cast<const uint4*>(raw);
    for (unsigned long long s0 = ((unsigned long long)blockIdx.x * 8 + wib) * 32; s0 < slots; s0 += (unsigned long long)gridDim.x * 8 * 32) {
        const uint32_t b = (uint32_t)(s0 / kRawBlock), e = (uint32_t)(s0 % kRawBlock) + lane;     // kRawBlock % 32 == 0: same block
        const bool live = e < __ldg(blk_count + b);
        uint4 x = make_uint4(0u, 0u, kAllNegative, kAllNegative), y = make_uint4(0u, kAllNegative, kAllNegative, 0u);
        if (live) { x = __ldg(ent + 2 * (s0 + lane)); y = __ldg(ent + 2 * (s0 + lane) + 1); }
        const bool acc16 = !(x.y & kRawFp32Flag);                 // the tile's accumulator type travels in the column words
        uint32_t z[4] = {~decode_sign_word(acc16, x.z), ~decode_sign_word(acc16, x.w), ~decode_sign_word(acc16, y.y), ~decode_sign_word(acc16, y.z)};
        const uint32_t first[2] = {x.y & ~kRawFp32Flag, y.x & ~kRawFp32Flag};
        const uint32_t c = __popc(z[0]) + __popc(z[1]) + __popc(z[2]) + __popc(z[3]);
        uint32_t incl = c;                                   // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += u; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        if (n + total > kStage) flush();
        Cand cd; cd.pos = x.x;
        if (total <= kStage) {
            uint32_t o = n + incl - c;
#pragma unroll
            for (int q = 0; q < 4; q++)
                while (z[q]) { const uint32_t bit = __ffs(z[q]) - 1; z[q] &= z[q] - 1; cd.col = column(acc16, first[q >> 1], q & 1, bit); st[o++] = cd; }
            n += total;
            if (n > 256) flush();
        } else {                                             // more than 20 candidates per entry on average: straight to global
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_cand, (unsigned long long)total);
            unsigned long long o = __shfl_sync(0xffffffffu, base, 0) + incl - c;
#pragma unroll
            for (int q = 0; q < 4; q++)
                while (z[q]) { const uint32_t bit = __ffs(z[q]) - 1; z[q] &= z[q] - 1; cd.col = column(acc16, first[q >> 1], q & 1, bit); if (o < cand_cap) cand[o] = cd; o++; }
        }
    }
    if (n) flush();
}

__global__ void __launch_bounds__(256)
rescore_kernel(MotifDev md, BlockDev blk, const Cand* __restrict__ cand,
               const unsigned long long* __restrict__ n_cand_ptr, unsigned long long cand_cap, HitSink sink)
{
    const bool masked = __ldg(blk.has_zero) != 0;   // block with zero-contribution characters: such positions add exactly 0
    unsigned long long n_cand = *n_cand_ptr;
    if (n_cand > cand_cap) n_cand = cand_cap;       // overflow: the host re-runs with a larger buffer
    // Hits of kRounds consecutive rounds are staged per warp in shared memory and appended with ONE global atomic
    // (all hits of a block go through a single counter: one atomic per warp round made the atomic unit the bottleneck).
    constexpr uint32_t kRounds = 4;
    __shared__ b200scan_hit s_hits[8][32 * kRounds];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    b200scan_hit* st = s_hits[wib];
    uint32_t n_st = 0, round = 0;
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
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < n_cand; base += stride) {
        unsigned long long i = base + threadIdx.x;
        bool hit = false;
        uint32_t pos = 0, col = 0;
        float s = 0.0f;
        if (i < n_cand && cand[i].col < md.n_cols) {      // (a candidate can never name a padding column; belt and braces)
            Cand c = cand[i];
            pos = c.pos; col = c.col;
            const uint32_t L = __ldg(md.len + col);
            const float* wp = reinterpret_cast<const float*>(md.w + __ldg(md.woff + col));
            uint32_t codes[4], zm[2] = {0u, 0u};
            load_window_codes(blk.codes, pos, codes);
            if (masked) load_window_zmask(blk.zmask, pos, zm);
            // 16 positions at a time: all (predicated) weight loads first -- independent, so one L2 round trip per group
            // instead of one per position -- then the additions strictly in position order
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((uint32_t)(16 * q) < L) {
                    const uint32_t r = codes[q], n = min(16u, L - 16u * q);
                    float wv[16];
#pragma unroll
                    for (uint32_t t = 0; t < 16; t++)
                        wv[t] = (t < n) ? __ldg(wp + 4 * (16 * q + t) + ((r >> (2 * t)) & 3u)) : 0.0f;
#pragma unroll
                    for (uint32_t t = 0; t < 16; t++)
                        if (t < n && !((zm[q >> 1] >> (16 * (q & 1) + t)) & 1u)) s += wv[t];
                }
            }
            hit = (pos < blk.n_payload) && !(s < __ldg(md.thr + col));
            if (hit) hit = window_in_fragment(blk, pos, L);
            col = __ldg(md.orig + col);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            b200scan_hit h; h.pos = pos; h.col = col; h.score = s;
            st[n_st + __popc(m & ((1u << lane) - 1u))] = h;
        }
        n_st += __popc(m);
        if (++round == kRounds) flush();
    }
    if (n_st) flush();
}

} // namespace b200
