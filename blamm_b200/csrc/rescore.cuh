// Exact rescoring of the candidates that the tensor-core filter let through.  One thread per candidate:
// the score is re-summed in FP32 in position order from the FP32 weights (bit-identical to the reference's
// sgemm chain, see gather.cuh), compared with the exact threshold, checked against the fragment table and
// the payload limit, and appended to the hit list.  Rare work (about 1e-4 of all scores), L2-resident.
#pragma once
#include "common.cuh"
#include "filter_tc.cuh"

namespace b200 {

// Raw entries of the tensor-core filter -> (position, sorted column) candidates.  One warp per block of entries; lane k
// tests word k of an entry (FP32 accumulators: word k = column first + k; FP16 via .pack::16b: word k = columns
// first + 2k in the low half and first + 2k + 1 in the high half; candidate <=> sign bit clear).  Candidates are staged
// in shared memory and appended with one global atomic per >= 256 of them.
template <bool ACC16>
__global__ void __launch_bounds__(256)
expand_kernel(const uint32_t* __restrict__ raw, const uint32_t* __restrict__ blk_count, const unsigned int* __restrict__ n_blocks_ptr,
              uint32_t blk_cap, Cand* __restrict__ cand, unsigned long long* n_cand, unsigned long long cand_cap,
              const uint32_t* __restrict__ has_zero)
{
    __shared__ Cand s_stage[8][512];      // per-warp staging (a trip adds <= 256): one global atomic per >= 256 candidates
    if (__ldg(has_zero) != 0) return;
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
    const uint32_t lt = (1u << lane) - 1u;
    // Each warp takes 4 consecutive entry slots (block, e) per trip and issues all their loads before using any
    // (the kernel is latency bound otherwise); slots beyond a block's count are skipped.
    const unsigned long long slots = (unsigned long long)nb * kRawBlock;
    for (unsigned long long s0 = ((unsigned long long)blockIdx.x * 8 + wib) * 4; s0 < slots; s0 += (unsigned long long)gridDim.x * 8 * 4) {
        const uint32_t b = (uint32_t)(s0 / kRawBlock), e0 = (uint32_t)(s0 % kRawBlock);       // kRawBlock % 4 == 0: same block
        const uint32_t cnt = __ldg(blk_count + b);
        uint32_t w[4], pos[4], first[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t* ent = raw + (s0 + k) * kRawWords;
            const bool live = e0 + k < cnt;
            w[k] = live ? __ldg(ent + lane) : 0x80008000u;           // all-negative: no candidate
            pos[k] = live ? __ldg(ent + 32) : 0u;
            first[k] = live ? __ldg(ent + 33) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (e0 + k >= cnt) break;                                 // warp-uniform
            Cand cd; cd.pos = pos[k];
            if (ACC16) {
                const bool lo = !(w[k] & 0x8000u), hi = !(w[k] & 0x80000000u);
                const unsigned blo = __ballot_sync(0xffffffffu, lo), bhi = __ballot_sync(0xffffffffu, hi);
                if (lo) { cd.col = first[k] + 2 * lane;     st[n + __popc(blo & lt)] = cd; }
                if (hi) { cd.col = first[k] + 2 * lane + 1; st[n + __popc(blo) + __popc(bhi & lt)] = cd; }
                n += __popc(blo) + __popc(bhi);
            } else {
                const bool c = (int32_t)w[k] >= 0;
                const unsigned bb = __ballot_sync(0xffffffffu, c);
                if (c) { cd.col = first[k] + lane; st[n + __popc(bb & lt)] = cd; }
                n += __popc(bb);
            }
        }
        if (n > 256) flush();
    }
    if (n) flush();
}

__global__ void __launch_bounds__(256)
rescore_kernel(MotifDev md, BlockDev blk, const Cand* __restrict__ cand,
               const unsigned long long* __restrict__ n_cand_ptr, unsigned long long cand_cap, HitSink sink)
{
    if (__ldg(blk.has_zero) != 0) return;          // such blocks went through the gather kernel
    unsigned long long n_cand = *n_cand_ptr;
    if (n_cand > cand_cap) n_cand = cand_cap;       // overflow: the host re-runs with a larger buffer
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
            uint32_t codes[4];
            load_window_codes(blk.codes, pos, codes);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((uint32_t)(16 * q) < L) {
                    uint32_t r = codes[q];
                    const uint32_t n = min(16u, L - 16u * q);
                    for (uint32_t t = 0; t < n; t++) {
                        s += __ldg(wp + 4 * (16 * q + t) + (r & 3u));
                        r >>= 2;
                    }
                }
            }
            hit = (pos < blk.n_payload) && !(s < __ldg(md.thr + col));
            if (hit) hit = window_in_fragment(blk, pos, L);
            col = __ldg(md.orig + col);
        }
        emit_hits_warp(hit, pos, col, s, sink);
    }
}

} // namespace b200
