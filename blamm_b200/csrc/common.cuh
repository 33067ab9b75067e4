// Shared device-side types and helpers for the b200scan kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/b200scan.h"

namespace b200 {

constexpr int kMaxLen = B200SCAN_MAX_MOTIF_LEN;   // 64 positions = 128 bits of 2-bit codes

// Per-species motif data on the device.  Columns are sorted by length (stable); `orig` maps back to the
// caller's column index (the reference's column order, motif.cpp:439-449).
struct MotifDev {
    const float4*   w;      // FP32 weights, one float4 (A,C,G,T) per position, column c at w[woff[c] .. +len[c])
    const uint32_t* woff;
    const uint32_t* len;
    const float*    thr;
    const uint32_t* orig;
    uint32_t        n_cols;
};

// The block resident on the device.
struct BlockDev {
    const uint32_t* codes;      // 2 bit / character, 16 per word
    const uint32_t* zmask;      // 1 bit / character, 32 per word (valid only if *has_zero != 0)
    const uint32_t* frag;       // ascending fragment starts (block positions), n_frag entries
    const uint32_t* has_zero;   // device flag written by the pack kernel / the submit call
    uint32_t n_total;
    uint32_t n_payload;
    uint32_t n_frag;
};

struct HitSink {
    b200scan_hit*        hits;
    unsigned long long*  n_hits;   // keeps counting past `cap` so the host can size a retry exactly
    unsigned long long   cap;
    uint32_t             compact;  // 1: the list holds 12-byte b200scan_hit12 records
    uint32_t*            bucket_cnt; // B200SCAN_HITS_8 (order.cuh): hits per COARSE bucket of 2^kCoarseShift positions, else nullptr
};
constexpr uint32_t kCoarseShift = 12;   // order.cuh: the device orders the hits of 4096 window positions per CTA
// One more hit at block position `pos` for the ordering pass: lanes of the calling warp that share a coarse bucket add up first
// (hits of one warp come from neighbouring windows).  Every lane named in `mask` must call this.
__device__ __forceinline__ void count_hit_bucket(const HitSink& sink, uint32_t mask, uint32_t pos)
{
    const uint32_t bucket = pos >> kCoarseShift;
    const uint32_t peers = __match_any_sync(mask, bucket);
    if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(sink.bucket_cnt + bucket, (uint32_t)__popc(peers));
}
// record `idx` of the hit list in the sink's format
__device__ __forceinline__ void store_hit(const HitSink& sink, unsigned long long idx, const b200scan_hit& h)
{
    if (sink.compact) {
        uint32_t* d = reinterpret_cast<uint32_t*>(sink.hits) + 3 * idx;
        d[0] = (uint32_t)h.pos; d[1] = h.col; d[2] = __float_as_uint(h.score);
    } else sink.hits[idx] = h;
}

// Window [pos, pos+L) lies wholly inside one fragment?  (SeqBlock::getRemainingSeqLen, sequence.cpp:68-79,
// used as `m.size() > remSeqLen -> reject` in pwmscan.cpp:122-126.)
__device__ __forceinline__ bool window_in_fragment(const BlockDev& b, uint32_t pos, uint32_t L)
{
    uint32_t lo = 0, hi = b.n_frag;
    while (lo < hi) {                       // first fragment start strictly greater than pos
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(b.frag + mid) > pos) hi = mid; else lo = mid + 1;
    }
    uint32_t end = (lo < b.n_frag) ? __ldg(b.frag + lo) : b.n_total;
    return (uint64_t)pos + L <= end;
}

// Characters left in the fragment that contains pos (SeqBlock::getRemainingSeqLen, sequence.cpp:68-79).
__device__ __forceinline__ uint32_t fragment_remaining(const BlockDev& b, uint32_t pos)
{
    uint32_t lo = 0, hi = b.n_frag;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(b.frag + mid) > pos) hi = mid; else lo = mid + 1;
    }
    return ((lo < b.n_frag) ? __ldg(b.frag + lo) : b.n_total) - pos;
}

// Warp-aggregated append: one global atomic per warp per call (every lane of the warp must call this).
__device__ __forceinline__ void emit_hits_warp(bool pred, uint32_t pos, uint32_t col, float score,
                                               const HitSink& sink)
{
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(sink.n_hits, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const unsigned long long idx = base + __popc(m & ((1u << lane) - 1u));
    const bool stored = pred && idx < sink.cap;
    if (stored) {
        b200scan_hit h; h.pos = pos; h.col = col; h.score = score;
        store_hit(sink, idx, h);
    }
    if (sink.bucket_cnt) {
        const uint32_t sm = __ballot_sync(0xffffffffu, stored);
        if (stored) count_hit_bucket(sink, sm, pos);
    }
}

// 64 characters of 2-bit codes starting at block position pos, as 4 words (character j in bits 2*(j%16) of
// word j/16).  Reads 5 words; the codes allocation is padded so this never leaves the buffer.
__device__ __forceinline__ void load_window_codes(const uint32_t* __restrict__ codes, uint32_t pos, uint32_t out[4])
{
    const uint32_t* p = codes + (pos >> 4);
    uint32_t sh = (pos & 15u) * 2u;
    uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
    out[0] = __funnelshift_r(w0, w1, sh);
    out[1] = __funnelshift_r(w1, w2, sh);
    out[2] = __funnelshift_r(w2, w3, sh);
    out[3] = __funnelshift_r(w3, w4, sh);
}

// 64 zero-mask bits starting at pos, as 2 words.
__device__ __forceinline__ void load_window_zmask(const uint32_t* __restrict__ zmask, uint32_t pos, uint32_t out[2])
{
    const uint32_t* p = zmask + (pos >> 5);
    uint32_t sh = pos & 31u;
    uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    out[0] = __funnelshift_r(w0, w1, sh);
    out[1] = __funnelshift_r(w1, w2, sh);
}

} // namespace b200
