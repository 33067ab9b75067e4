// Device-side ordering of a block's hit list: (position, column) order and 8-byte records (B200SCAN_HITS_8).
//
// What it replaces: in the reference the hits of a block reach PWMScan::writeOccToDisk in the order the R matrix is swept
// (pwmscan.cpp:108-131, per offset: column-major); with threads the order of the occurrence file is unspecified
// (README.md:165).  The drop-in CLI writes every block in (position, column) order; until round 2 the host radix-sorted 12-byte
// records for that (half of its formatting time) and every hit crossed PCIe as 12 bytes -- the hit list IS the host-link traffic
// of this path (3 bytes of hits per character at -pt 1e-4 x 1800 columns).  Here the device does the ordering, in two levels:
//
//   rescore / gather   also count hits per COARSE bucket of 4096 window positions        (one warp-aggregated RED per group of lanes)
//   bucket_scan        exclusive scan of the coarse counts -> coarse_start[], cursor[]    (one CTA; 24 k counters per 100 Mbp)
//   bucket_scatter     unordered 12-byte list -> coarse-bucket-contiguous list           (one warp-aggregated ATOM per group of lanes)
//   bucket_order       one CTA per coarse bucket: counting sort by the position inside the bucket (4096 shared-memory bins), the
//                      positions that hold several hits are then put in column order (compacted list of such positions: one
//                      round for the whole CTA); writes the final 8-byte records
//                          key = (pos & 255) << 24 | column,  score
//                      and bucket_start[] of the FINE buckets of 256 positions the ABI hands to the host (prefix sums of the bins
//                      at multiples of 256): the position of hit i of fine bucket f is 256 f + (key >> 24).
//
// HBM / L2 bound (no arithmetic to speak of): per hit 12 B read + 12 B written (scatter) + 24 B read (two passes of the order kernel,
// the second one out of L2) + 8 B written.  All integer / byte work, bit-exact by construction; tests compare with the oracle's order.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr uint32_t kBucketShift = B200SCAN_BUCKET_SHIFT;          // ABI: 256 window positions per (fine) bucket
constexpr uint32_t kBucketSize  = 1u << kBucketShift;
constexpr uint32_t kCoarseSize  = 1u << kCoarseShift;             // 4096 positions per coarse bucket (kCoarseShift: common.cuh)
constexpr uint32_t kFinePerCoarse = kCoarseSize / kBucketSize;    // 16
static_assert(kBucketShift == 8, "hit8 keys carry 8 position bits and 24 column bits");

struct Hit12 { uint32_t pos, col, score; };

// Exclusive scan of cnt[0 .. n) into start[0 .. n] (start[n] = total) and cursor[i] = start[i]; one CTA of 1024 threads,
// 8 counters per thread per round.  n is at most 2^20 (a 2^32-character block): 128 rounds; a 100 Mbp block takes 3.
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

// Unordered list -> coarse-bucket-contiguous list (order inside a bucket arbitrary).  Consecutive records of the unordered list
// come from one warp's rescoring rounds, i.e. from neighbouring windows: the lanes of a warp mostly share one or two buckets and
// take their slots with one atomic per group (__match_any_sync).
__global__ void __launch_bounds__(256)
bucket_scatter_kernel(const Hit12* __restrict__ in, const unsigned long long* __restrict__ n_hits_ptr, unsigned long long cap,
                      uint32_t* __restrict__ cursor, Hit12* __restrict__ out)
{
    unsigned long long n = *n_hits_ptr;
    if (n > cap) n = cap;                                            // overflow: the host re-runs with larger buffers
    const uint32_t lane = threadIdx.x & 31;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n; i0 += stride) {
        const unsigned long long i = i0 + lane;
        const bool valid = i < n;
        Hit12 h{0, 0, 0};
        if (valid) h = in[i];
        const uint32_t bucket = valid ? (h.pos >> kCoarseShift) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) base = atomicAdd(cursor + bucket, (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) out[base + __popc(peers & ((1u << lane) - 1u))] = h;
    }
}

// One CTA per coarse bucket (grid-stride).  Pass 1 counts the bucket's hits per position (4096 shared-memory bins), a block scan
// turns the bins into offsets (and yields bucket_start of the 16 fine buckets), pass 2 places every hit at offset[position]++ as
// an 8-byte record; the positions that hold several hits (a few per cent at the usual densities) are collected in a list and put
// in column order, one thread per position (selection sort in place: the runs are tiny).
// Usual case (<= kStageHits hits in the bucket, about 1000 at -pt 1e-4 x 1800 columns): every thread loads its <= 8 hits ONCE, all
// loads in flight together (the list was just written by the scatter pass and mostly lies in DRAM: one latency instead of eight),
// keeps them in registers across the two passes, the records are placed and tie-ordered in shared memory and leave with coalesced
// stores.  Denser buckets take the same steps out of global memory.
constexpr uint32_t kOrderThreads = 256;
constexpr uint32_t kBinsPerThread = kCoarseSize / kOrderThreads;  // 16
constexpr uint32_t kHitsPerThread = 8;
constexpr uint32_t kStageHits = kOrderThreads * kHitsPerThread;   // 2048 records = 16 KB of shared memory
constexpr uint32_t kMultiCap = 512;

// bin of position p: the scan reads bins 16 t .. 16 t + 15 from thread t -- XOR with the thread's low bits spreads them over the banks
__device__ __forceinline__ uint32_t bin_at(uint32_t p) { return p ^ ((p >> 4) & 15u); }

template <class T> __device__ __forceinline__ void order_run(T* r, uint32_t m)
{
    for (uint32_t a = 0; a + 1 < m; a++) {                           // keys of one position differ only in the column bits
        uint32_t best = a; uint2 vb = r[a];
        for (uint32_t q = a + 1; q < m; q++) { const uint2 vq = r[q]; if (vq.x < vb.x) { best = q; vb = vq; } }
        if (best != a) { r[best] = r[a]; r[a] = vb; }
    }
}

__global__ void __launch_bounds__(kOrderThreads)
bucket_order_kernel(const Hit12* __restrict__ in, const uint32_t* __restrict__ coarse_start, uint32_t n_coarse, uint32_t n_fine,
                    uint32_t* __restrict__ bucket_start, uint2* __restrict__ out)
{
    __shared__ uint32_t s_bins[kCoarseSize];
    __shared__ uint2    s_rec[kStageHits];
    __shared__ uint32_t s_warp[kOrderThreads / 32];
    __shared__ uint32_t s_multi[kMultiCap];                          // first slot (relative to the bucket) of positions with several hits
    __shared__ uint32_t s_mlen[kMultiCap];
    __shared__ uint32_t s_nmulti;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    for (uint32_t b = blockIdx.x; b < n_coarse; b += gridDim.x) {
        const uint32_t lo = coarse_start[b], hi = coarse_start[b + 1];
        if (hi == lo) {                                              // empty: only the fine bucket index
            if (tid <= kFinePerCoarse) {
                const uint32_t f = b * kFinePerCoarse + tid;
                if ((tid < kFinePerCoarse && f < n_fine) || (b == n_coarse - 1 && f == n_fine)) bucket_start[f] = lo;
            }
            continue;
        }
        const bool staged = hi - lo <= kStageHits;
        Hit12 h[kHitsPerThread];
        if (staged) {
#pragma unroll
            for (uint32_t r = 0; r < kHitsPerThread; r++) {
                const uint32_t i = lo + tid + kOrderThreads * r;
                if (lo + (tid & ~31u) + kOrderThreads * r >= hi) break;        // (warp-uniform: the usual bucket holds ~4 hits per thread, not 8)
                if (i < hi) h[r] = in[i];
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kBinsPerThread; k++) s_bins[tid + kOrderThreads * k] = 0u;
        if (tid == 0) s_nmulti = 0;
        __syncthreads();
        if (staged) {
#pragma unroll
            for (uint32_t r = 0; r < kHitsPerThread; r++) {
                if (lo + (tid & ~31u) + kOrderThreads * r >= hi) break;
                if (lo + tid + kOrderThreads * r < hi) atomicAdd(s_bins + bin_at(h[r].pos & (kCoarseSize - 1)), 1u);
            }
        } else {
            for (uint32_t i = lo + tid; i < hi; i += kOrderThreads) atomicAdd(s_bins + bin_at(in[i].pos & (kCoarseSize - 1)), 1u);
        }
        __syncthreads();
        // exclusive scan over the 4096 bins: thread t owns the positions 16 t .. 16 t + 15
        uint32_t c[kBinsPerThread], sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < kBinsPerThread; k++) { c[k] = s_bins[bin_at(kBinsPerThread * tid + k)]; sum += c[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += u; }
        if (lane == 31) s_warp[wib] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (uint32_t w = 0; w < wib; w++) run += s_warp[w];
        // fine bucket j of this coarse bucket starts at the offset of thread 16 j
        if ((tid & (kFinePerCoarse - 1)) == 0) {
            const uint32_t f = b * kFinePerCoarse + tid / kFinePerCoarse;
            if (f < n_fine) bucket_start[f] = lo + run;
        }
        if (b == n_coarse - 1 && tid == 0) bucket_start[n_fine] = hi;
        const uint32_t first = run;
#pragma unroll
        for (uint32_t k = 0; k < kBinsPerThread; k++) {
            s_bins[bin_at(kBinsPerThread * tid + k)] = run;
            if (c[k] >= 2) {
                const uint32_t slot = atomicAdd(&s_nmulti, 1u);
                if (slot < kMultiCap) { s_multi[slot] = run; s_mlen[slot] = c[k]; }
            }
            run += c[k];
        }
        __syncthreads();
        if (staged) {
#pragma unroll
            for (uint32_t r = 0; r < kHitsPerThread; r++) {
                if (lo + (tid & ~31u) + kOrderThreads * r >= hi) break;
                if (lo + tid + kOrderThreads * r < hi) {
                    const uint32_t p = h[r].pos & (kCoarseSize - 1);
                    s_rec[atomicAdd(s_bins + bin_at(p), 1u)] = make_uint2(((p & (kBucketSize - 1)) << 24) | h[r].col, h[r].score);
                }
            }
        } else {
            for (uint32_t i = lo + tid; i < hi; i += kOrderThreads) {
                const Hit12 g = in[i];
                const uint32_t p = g.pos & (kCoarseSize - 1);
                out[lo + atomicAdd(s_bins + bin_at(p), 1u)] = make_uint2(((p & (kBucketSize - 1)) << 24) | g.col, g.score);
            }
        }
        __syncthreads();                                             // the records of this bucket are visible to the whole CTA
        const uint32_t nm = s_nmulti;
        if (nm <= kMultiCap) {
            for (uint32_t j = tid; j < nm; j += kOrderThreads) { if (staged) order_run(s_rec + s_multi[j], s_mlen[j]); else order_run(out + lo + s_multi[j], s_mlen[j]); }
        } else {                                                     // many ties: every thread orders the runs of its own positions
            uint32_t at = first;
#pragma unroll
            for (uint32_t k = 0; k < kBinsPerThread; k++) {
                if (c[k] >= 2) { if (staged) order_run(s_rec + at, c[k]); else order_run(out + lo + at, c[k]); }
                at += c[k];
            }
        }
        __syncthreads();
        if (staged) {
            for (uint32_t i = tid; i < hi - lo; i += kOrderThreads) out[lo + i] = s_rec[i];
            __syncthreads();                                         // s_rec / s_bins / s_nmulti are reused by the next bucket
        }
    }
}

} // namespace b200
