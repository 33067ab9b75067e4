// Exact rescoring of the candidates that the tensor-core filter let through.  One thread per candidate:
// the score is re-summed in FP32 in position order from the FP32 weights (bit-identical to the reference's
// sgemm chain, see gather.cuh), compared with the exact threshold, checked against the fragment table and
// the payload limit, and appended to the hit list.  Rare work (about 1e-4 of all scores), L2-resident.
#pragma once
#include "common.cuh"

namespace b200 {

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
        if (i < n_cand) {
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
