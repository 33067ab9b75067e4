// Device-side ordering of a block's hit list: (position, column) order and 8-byte records (B200SCAN_HITS_8).
//
// What it replaces: in the reference the hits of a block reach PWMScan::writeOccToDisk in the order the R matrix is swept
// (pwmscan.cpp:108-131, per offset: column-major); with threads the order of the occurrence file is unspecified
// (README.md:165).  The drop-in CLI writes every block in (position, column) order; until round 2 the host radix-sorted 12-byte
// records for that (half of its formatting time) and every hit crossed PCIe as 12 bytes -- the hit list IS the host-link traffic
// of this path (3 bytes of hits per character at -pt 1e-4 x 1800 columns).  Here the device does the ordering:
//
//   rescore_kernel   (rescore.cuh) also counts hits per BUCKET of 256 window positions          (one RED per hit)
//   bucket_scan      exclusive scan of the counts -> bucket_start[0 .. n_buckets]                (one CTA)
//   bucket_scatter   unordered 12-byte list -> bucket-contiguous list                           (one ATOM per hit)
//   bucket_order     one warp per bucket: counting sort by the position inside the bucket (256 shared-memory bins), then the few
//                    positions that hold several hits are put in column order; writes the final 8-byte records
//                        key = (pos & 255) << 24 | column,  score
//                    The host gets the records plus bucket_start: the position of hit i of bucket b is 256 b + (key >> 24).
//
// HBM bound (no arithmetic to speak of): per hit 12 B read + 12 B written (scatter) + 24 B read (two passes of the order kernel,
// the second one out of L2) + 8 B written.  All integer / byte work, bit-exact by construction; tests compare with the host sort.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t kBucketShift = B200SCAN_BUCKET_SHIFT;          // 256 window positions per bucket
constexpr uint32_t kBucketSize  = 1u << kBucketShift;
static_assert(kBucketShift == 8, "hit8 keys carry 8 position bits and 24 column bits");

struct Hit12 { uint32_t pos, col, score; };

// Exclusive scan of cnt[0 .. n) into start[0 .. n] (start[n] = total) and cursor[i] = start[i]; one CTA of 1024 threads,
// 8 counters per thread per round.  n is at most 2^24 (a 2^32-character block): 2048 rounds; typical blocks take 16-50.
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(const uint32_t* __restrict__ cnt, uint32_t n, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 8192) {
        const uint32_t i0 = base + threadIdx.x * 8;
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = (i0 + k < n) ? cnt[i0 + k] : 0u; sum += v[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += u; }
        if (lane == 31) s_warp[wib] = incl;
        __syncthreads();
        if (wib == 0) {
            uint32_t w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, d); if ((int)lane >= d) wi += u; }
            s_warp[lane] = wi - w;                                   // exclusive prefix of the warp sums
        }
        __syncthreads();
        uint32_t run = s_carry + s_warp[wib] + incl - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (i0 + k < n) { start[i0 + k] = run; cursor[i0 + k] = run; }
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[n] = s_carry;
}

// Unordered list -> bucket-contiguous list (order inside a bucket arbitrary).
__global__ void __launch_bounds__(256)
bucket_scatter_kernel(const Hit12* __restrict__ in, const unsigned long long* __restrict__ n_hits_ptr, unsigned long long cap,
                      uint32_t* __restrict__ cursor, Hit12* __restrict__ out)
{
    unsigned long long n = *n_hits_ptr;
    if (n > cap) n = cap;                                            // overflow: the host re-runs with larger buffers
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Hit12 h = in[i];
        const uint32_t slot = atomicAdd(cursor + (h.pos >> kBucketShift), 1u);
        out[slot] = h;
    }
}

// One warp per bucket.  Pass 1 counts the bucket's hits per position (shared-memory bins), a warp scan turns the bins into
// offsets, pass 2 places every hit at start + offset[position]++ as an 8-byte record; positions with several hits (a few per
// bucket at the usual densities) are then put in column order by one lane each (selection sort in place: the run is tiny).
__global__ void __launch_bounds__(256)
bucket_order_kernel(const Hit12* __restrict__ in, const uint32_t* __restrict__ start, uint32_t n_buckets, uint2* __restrict__ out)
{
    __shared__ uint32_t s_bins[8][kBucketSize];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* bins = s_bins[wib];
    for (uint32_t b = blockIdx.x * 8 + wib; b < n_buckets; b += gridDim.x * 8) {
        const uint32_t lo = __ldg(start + b), hi = __ldg(start + b + 1);
        if (hi == lo) continue;
        if (hi - lo == 1) {                                          // a lone hit needs no ordering
            if (lane == 0) { const Hit12 h = in[lo]; out[lo] = make_uint2(((h.pos & (kBucketSize - 1)) << 24) | h.col, h.score); }
            continue;
        }
#pragma unroll
        for (uint32_t k = 0; k < kBucketSize / 32; k++) bins[lane + 32 * k] = 0u;
        __syncwarp();
        for (uint32_t i = lo + lane; i < hi; i += 32) atomicAdd(bins + (in[i].pos & (kBucketSize - 1)), 1u);
        __syncwarp();
        // exclusive scan over the 256 bins: lane l owns bins 8 l .. 8 l + 7
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { c[k] = bins[8 * lane + k]; sum += c[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += u; }
        uint32_t run = lo + incl - sum;
        uint32_t first[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { first[k] = run; bins[8 * lane + k] = run; run += c[k]; }
        __syncwarp();
        for (uint32_t i = lo + lane; i < hi; i += 32) {
            const Hit12 h = in[i];
            const uint32_t p = h.pos & (kBucketSize - 1);
            const uint32_t slot = atomicAdd(bins + p, 1u);
            out[slot] = make_uint2((p << 24) | h.col, h.score);
        }
        __syncwarp();                                                // the records of this bucket are visible to the whole warp
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (c[k] < 2) continue;
            uint2* r = out + first[k];
            for (uint32_t a = 0; a + 1 < c[k]; a++) {                // keys of one position differ only in the column bits
                uint32_t m = a; uint2 vm = r[a];
                for (uint32_t q = a + 1; q < c[k]; q++) { const uint2 vq = r[q]; if (vq.x < vm.x) { m = q; vm = vq; } }
                if (m != a) { r[m] = r[a]; r[a] = vm; }
            }
        }
        __syncwarp();
    }
}

} // namespace b200
