// tcgen05 / TMEM filter kernel: the motif x window contraction on the 5th-gen tensor cores.
//
// What it replaces: the reference's  R = S[:, 4*off : 4*off+K] * P  (cublasSgemm per offset per tile,
// matrix.h:314-323, pwmscan.cpp:385-389) followed by filterScore (kernel.cu:21-34).  Differences by design:
//
//  * The window operand is never materialised in HBM.  2-bit codes are bulk-copied (TMA, cp.async.bulk ->
//    UBLKCP) into shared memory and expanded on chip into a linear array E of 16-byte entries
//        E[p] = [ onehot_f16(code[p]) (4 halves) | onehot_f16(code[p+1]) (4 halves) ].
//    With the no-swizzle K-major canonical layout, a UMMA operand row is 16 B and consecutive rows are 16 B
//    apart, so "row r, K-chunk kk" of the Toeplitz window matrix is simply E[g0 + r + 2*kk]: the shared
//    memory descriptor (LBO = 32 B between K-chunks, SBO = 128 B between 8-row groups) reads the
//    overlapping windows straight out of E.  128 windows x (4 positions per MMA) cost 2 KB of smem, not
//    128 x K x 2 B.
//  * All columns of a length-bucket tile (<= 256 columns, FP16 weights, threshold folded in) stay resident
//    in shared memory; a CTA sweeps windows, D[128 windows x N columns] accumulates in TMEM, double buffered
//    so the epilogue of tile i overlaps the MMAs of tile i+1.  Accumulators are FP16 when the host's
//    per-column error bound allows it (tcgen05.ld .pack::16b then delivers two of them per register, halving
//    the epilogue's ALU work), else FP32.
//  * The tensor pass is a CONSERVATIVE FILTER: weights are  fp16_round_up(W[j][b] - thr'/L)  with
//    thr' = thr - margin, so  acc >= 0  whenever the exact FP32 score >= thr (b200scan.cu: build_tc_tiles).
//    The epilogue only looks at sign bits (one PRMT + one IMAD per four scores) and appends raw entries;
//    rescore.cuh expands them to (pos, col) candidates and recomputes those few scores exactly.  R never exists.
//  * Blocks with zero-contribution characters (lower case under the reference's BLAS-path semantics) run the ZMASK = true
//    instance: masked characters become all-zero operand rows, the weights stay unshifted and one leading MMA step adds the
//    bias -(thr - margin) through a constant one-hot operand (b200scan.cu: fold_z).
//
// Roofline: tensor pipe.  One tcgen05.mma (M=128, N, K=16) covers 4 motif positions of N columns for 128
// windows and takes N/2 cycles; the epilogue must drain 128 x N accumulators per tile from TMEM.
//
// Warp roles (608 threads, 1 CTA/SM, persistent with an atomic work counter):
//   warps 0..1   producers: codes -> E ring (3 stages of 4 tiles = 512 entries + mirrored halo) through a 16-entry one-hot LUT
//   warp 2       TMEM allocation; ONE elected lane runs the whole issue loop: tcgen05.mma chains and tcgen05.commit, one
//                eFull wait / eEmpty commit per stage, one tEmpty wait / tFull commit per tile (B/codes bulk loads: thread 0)
//   warps 3..18  epilogue: TWO independent groups of 8 warps, group g owns TMEM buffer g (even / odd window tiles).
//                Within a group two warps share each TMEM lane quarter and split the tile's 32-word chunks:
//                tcgen05.ld 32x32b.x32 into registers, RELEASE the TMEM buffer at once, compact the sign bits of the 32
//                words into two sign words (PRMT on the ALU pipe + IMAD on the FMA pipe).  A lane that saw a non-negative
//                accumulator stores {window, column, sign words} as one 32-byte raw entry in global memory (blocks of 64
//                entries reserved ahead of time, one predicated fire-and-forget store).  Because a buffer is released
//                right after the load, candidate pushes land in slack instead of delaying the next MMAs.
//   rescore_tile_kernel (rescore.cuh) later decodes the raw entries and recomputes those few scores exactly, a column tile at a time.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace b200 {

#ifndef TC_EPI_WARPS
#define TC_EPI_WARPS 16
#endif
#ifndef TC_KNOCKOUT
#define TC_KNOCKOUT 0      // diagnostic only: 1 skip epilogue ld+reduce, 2 skip MMA issue, 4 skip producer fill, 8 skip FIFO push
#endif
#ifndef TC_EPI_GROUPS
#define TC_EPI_GROUPS 2       // 2: two independent groups of epilogue warps, group g owns TMEM buffer g (even / odd tiles)
#endif
#ifndef TC_PRODUCERS
#define TC_PRODUCERS 2
#endif
#ifndef TC_MMA_WARPS
#define TC_MMA_WARPS 1        // issuer warps; 2 (experimental, single-CTA instances only): issuer m takes the tiles whose running index is m mod 2
#endif
#ifndef TC_STAGE_TILES
#define TC_STAGE_TILES 4      // 128-window tiles per E stage: the issuer pays one eFull wait and one eEmpty commit per stage
#endif
#ifndef TC_BUFS
#define TC_BUFS 2
#endif
#ifndef TC_MAXN
#define TC_MAXN 256
#endif
#ifndef TC_CTAS_PER_SM
#define TC_CTAS_PER_SM 1
#endif
#ifndef TC_ISSUE_FAST
#define TC_ISSUE_FAST 1       // 1: single-CTA instances issue whole E stages (4 tiles) from an unrolled loop with loop-invariant operands
#endif
#ifndef TC_HALF_TAGS
#define TC_HALF_TAGS 0        // 1: the specialised epilogue tags its raw blocks per 128-column half (more, smaller weight tables for the fused rescorer)
#endif
#ifndef TC_EPI_FAST
#define TC_EPI_FAST 1         // 1: 256-column tiles with packed accumulators take the specialised epilogue loop (one LDTM.x64 per warp at a precomputed
#endif                        //    address, maximum pre-test, sign compaction only for the 16-word groups that hold a candidate)
constexpr uint32_t kTcEpiWarps = TC_EPI_WARPS;                 // per TMEM lane quarter: kTcEpiWarps / 4
constexpr uint32_t kTcCtasPerSm = TC_CTAS_PER_SM;             // co-resident CTAs share the SM's 512 TMEM columns
constexpr uint32_t kTcProducers = TC_PRODUCERS;                         // warps filling the E ring, 32 entries of every stage each
constexpr uint32_t kTcMmaWarps = TC_MMA_WARPS;
constexpr uint32_t kTcEpiWarp0 = kTcProducers + kTcMmaWarps;  // first epilogue warp (warps kTcProducers .. issue the MMAs)
constexpr int      kTcThreads  = 32 * (kTcProducers + kTcMmaWarps + kTcEpiWarps);
static_assert(TC_MMA_WARPS == 1 || TC_MMA_WARPS == 2, "one or two issuer warps");
#ifndef TC_SPAN
#define TC_SPAN 65536      // (32768: 1.4 % slower on the bench set -- every item pays one pipeline fill and drain)
#endif
constexpr uint32_t kTcSpan     = TC_SPAN;    // windows per work item
constexpr uint32_t kTcStageTiles = TC_STAGE_TILES;
constexpr uint32_t kTcStageEnt = 128 * kTcStageTiles;          // entries per E stage
constexpr uint32_t kTcStages   = kTcStageTiles >= 4 ? 4 : 8 / kTcStageTiles;        // E ring stages: a power of two <= 8 (barrier map), ring of >= 8 tiles
static_assert((kTcStages & (kTcStages - 1)) == 0 && (TC_BUFS & (TC_BUFS - 1)) == 0, "ring and buffer counts are powers of two (the issuer steps them with masks)");
constexpr uint32_t kTcMirror   = 64;         // entries mirrored past the ring end (>= 2*(2*nK_max-1))
constexpr uint32_t kTcMaxN     = TC_MAXN;     // columns per tile; 2 accumulator buffers of kTcMaxN TMEM columns per CTA
static_assert(TC_BUFS * TC_MAXN * TC_CTAS_PER_SM <= 512, "TMEM: buffers x N columns x CTAs per SM must fit 512 columns");
static_assert((4 * TC_STAGE_TILES) % TC_PRODUCERS == 0 && TC_STAGE_TILES >= 1 && TC_STAGE_TILES <= 8 && TC_EPI_WARPS % (4 * TC_EPI_GROUPS) == 0 && (TC_EPI_GROUPS & (TC_EPI_GROUPS - 1)) == 0 && TC_BUFS % TC_EPI_GROUPS == 0, "warp role split");
constexpr uint32_t kTcEpiGroups = TC_EPI_GROUPS;
constexpr uint32_t kRawBlock   = 64;         // raw entries per block (an epilogue warp reserves a block at a time)
constexpr uint32_t kRawWords   = 8;          // {window, column of chunk 0, 2 sign words, column of chunk 1, 2 sign words, -} = 32 B per entry

struct TcTile {
    uint32_t col0;      // first sorted column
    uint32_t n_cols;    // real columns
    uint32_t n_pad;     // N of the MMA, multiple of 64, <= 256 (padding columns can never pass the filter)
    uint32_t n_k;       // MMAs per 128-window tile = ceil(Lmax / 4)
    uint32_t b_off;     // byte offset of the tile's B image in TcParams::bimg
    uint32_t b_bytes;   // n_pad * (2*n_k) * 16
    uint32_t acc16;     // 1: FP16 accumulators (filter_tc_kernel<true>), 0: FP32 -- chosen per tile by the host's error bound; the host
                        // launches one kernel instance per accumulator type over its share of the tiles
};

struct TcParams {
    const uint8_t* bimg;
    const TcTile*  tiles;
    uint32_t       n_tiles;
    uint32_t       n_spans;
    unsigned int*  work_counter;
    uint32_t*      raw;            // raw entries: block b, entry e at raw[(b * kRawBlock + e) * kRawWords]
    uint32_t*      blk_count;      // entries used in block b
    uint32_t*      blk_tag;        // column tile of block b: first sorted column << 9 | real columns (rescore.cuh: rescore_tile_kernel)
    unsigned int*  n_blocks;       // blocks reserved so far (keeps counting past blk_cap so the host can size a retry)
    uint32_t       blk_cap;
    unsigned int*  error_flag;
    unsigned long long* trace;      // B200_TRACE builds only: [role 0..3][tile 0..kTraceTiles)[event 0..3] clock64 stamps of CTA 0, item 0
};

constexpr uint32_t kTraceTiles = 256;
#if defined(B200_TRACE) && !defined(B200_PHASE)
#define TC_TRACE(role, tile_i, ev) do { if (blockIdx.x == 0 && nItem == 0 && lane == 0 && (tile_i) < kTraceTiles) \
        P.trace[((role) * kTraceTiles + (tile_i)) * 4 + (ev)] = (unsigned long long)clock64(); } while (0)
#else
#define TC_TRACE(role, tile_i, ev) do {} while (0)
#endif
// B200_PHASE builds (diagnostic, tools/tc_phase.py): per-role cycle totals over the whole launch, accumulated in registers
// (two clock reads per phase, no memory traffic until the kernel ends).  P.trace[16 * role + k], summed over all CTAs:
//   role 0 producer warp 0:   0 waiting for a free E stage   1 filling            3 stages
//   role 1 issuer:            0 waiting for E stages         1 waiting for TMEM   2 issuing + commits   3 tiles
//   role 2 epilogue, group 0: 0 waiting for tFull            1 tcgen05.ld phase   2 compaction + push   3 tiles
#ifdef B200_PHASE
#define PH_MARK()   do { ph_t = clock64(); } while (0)
#define PH_ACC(k)   do { const long long n_ = clock64(); ph_a[k] += n_ - ph_t; ph_t = n_; } while (0)
#define PH_COUNT()  do { ph_a[3]++; } while (0)
#else
#define PH_MARK()   do {} while (0)
#define PH_ACC(k)   do {} while (0)
#define PH_COUNT()  do {} while (0)
#endif

// shared memory carve-up (bytes)
constexpr uint32_t kSmE      = (kTcStages * kTcStageEnt + kTcMirror) * 16;
constexpr uint32_t kSmCodes  = kTcSpan / 4 + 128;                              // 16512
constexpr uint32_t kSmB      = kTcMaxN * (2 * (kMaxLen / 4 + 1)) * 16;         // 139264: up to 16 position steps + the bias step of masked blocks
constexpr uint32_t kSmZ      = kTcSpan / 8 + 64;                               //   8256: zero-mask bits of the span (masked blocks only)
constexpr uint32_t kSmOnes   = 144 * 16;                                       //   2304: constant operand of the bias step
constexpr uint32_t kSmBars   = 32 * 8;
constexpr uint32_t kSmLut    = 16 * 16;                                        //   256: E entry for every (code, next code)
constexpr uint32_t kTcSmemBytes = kSmE + kSmCodes + kSmB + kSmBars + kSmLut + kSmZ + kSmOnes + 128;

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must trap (launch failure), never hang the device.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, unsigned int* error_flag) {
    const long long t0 = clock64();
#pragma unroll 1
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {            // ~2 s at 2 GHz: orders of magnitude beyond any legal wait
            atomicExch(error_flag, 0xDEAD0000u | (bar & 0xFFFFu));
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* error_flag) {
    if (mbar_try(bar, parity)) return;
    mbar_wait_slow(bar, parity, error_flag);
}
// The same wait without a function call: the slow path above is a CALL, and across a call ptxas gives up everything it holds in
// uniform registers (descriptors, TMEM and barrier addresses are then re-derived and moved with R2UR after every wake-up -- on
// the issuing lane that is the critical path of every tile).  Bounded by a spin count instead of the clock (every failed
// mbarrier.try_wait already suspends the thread for a hardware time slice, so 2^26 of them are far beyond any legal wait).
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity, unsigned int* error_flag) {
    uint32_t spins = 0;
#pragma unroll 1
    while (!mbar_try(bar, parity)) {
        if (++spins > (1u << 26)) { atomicExch(error_flag, 0xDEAD0000u | (bar & 0xFFFFu)); __trap(); }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init()   { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before()   { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()    { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One elected lane of a fully converged warp (elect.sync).  Keeping the issuing warp converged and predicating only the
// tcgen05 instruction lets ptxas keep descriptors in uniform registers; an `if (lane == 0)` block instead costs an
// ELECT + R2UR.BROADCAST + BRA.U.ANY loop (~100 cycles) per instruction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
// Same, with the descriptors as (lo, hi) words: stepping along K is then a single 32-bit add on the address field.
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}"
                 ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
// Same shape with .pack::16b: register k = the low 16 bits of TMEM columns 2k (low half) and 2k+1 (high half), i.e.
// 64 FP16 accumulators (stored one per 32-bit column) arrive as 32 registers.
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
// 64 registers in one instruction (LDTM.x64): with .pack::16b that is 128 TMEM columns = half of a 256-column tile per warp
__device__ __forceinline__ void tmem_ld64_pack16(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr) : "memory");
}
template <bool ACC16> __device__ __forceinline__ void tmem_ld_words(uint32_t taddr_buf, uint32_t word, uint32_t (&v)[32]) {
    if (ACC16) tmem_ld32_pack16(taddr_buf + 2 * word, v); else tmem_ld32(taddr_buf + word, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: one issuer drives the tensor cores of two SMs (M = 256: each CTA supplies its own 128
// window rows and half of the B columns); completion is multicast to the barriers of both CTAs; the peer CTA signals the leader's
// barriers through the cluster shared window.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {        // address of the same smem location in CTA `rank`
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t raddr) {           // cluster-scope release: orders this CTA's smem writes
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_cta(uint32_t raddr) {       // default (CTA-scope) semantics: no cluster-wide fence; enough when
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");   // only tcgen05 operations are being ordered (TMEM hand-back)
}
__device__ __forceinline__ void st_cluster_u32(uint32_t raddr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_f16_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}"
                 ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {          // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// kind::i8 (UTCIMMA): signed 8-bit operands, K = 32 per instruction, S32 accumulators
__device__ __forceinline__ void umma_i8_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}"
                 ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool PAIR, bool I8> __device__ __forceinline__ void umma_x(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
    if (I8) umma_i8_lohi(d, alo, ahi, blo, bhi, idesc, acc);
    else if (PAIR) umma2_f16_lohi(d, alo, ahi, blo, bhi, idesc, acc); else umma_f16_lohi(d, alo, ahi, blo, bhi, idesc, acc);
}
template <bool PAIR> __device__ __forceinline__ void commit_x(uint32_t bar) { if (PAIR) umma2_commit(bar); else umma_commit(bar); }

// K-major, no-swizzle ("interleave") shared memory descriptor: rows 16 B apart inside an 8-row core matrix,
// 8-row groups SBO apart, the two 16-byte K-chunks of one K=16 MMA LBO apart (cute mma_traits_sm100.hpp,
// canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; version = 1 on sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}


// Sign bits of the 32 words one tcgen05.ld delivered, compacted into two words.  PRMT in sign-replication mode (ALU pipe)
// turns the signs of two words into 0x00 / 0xFF bytes r_t; an IMAD (FMA pipe, otherwise idle here) accumulates
// X = sum_t r_t * 2^t.  With M_b = sum_t 2^t [accumulator (t, b) negative] that is X = sum_b 255 * M_b * 256^b (mod 2^32),
// which rescore.cuh: decode_sign_word() inverts; all negative <=> X == kAllNegative.  (An AND/OR accumulation would put
// all 34 operations per chunk on the half-rate ALU pipe, which the epilogue warps then saturate.)
//   FP16 accumulators (.pack::16b: word k = columns 2k | 2k+1): 64 columns per chunk; (t, b) of word w <-> column 32w + 4t + b
//   FP32 accumulators (word k = column k):                      32 columns per chunk; (t, b) of word w <-> column 16w + 2t + b, b < 2
constexpr uint32_t kAllNegative = 0xFFFFFF01u;
constexpr uint32_t kRawFp32Flag = 0x80000000u;     // set in a raw entry's column words when its tile used FP32 accumulators
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {         // a * b + c, kept a multiply-add
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <bool ACC16, int OFF, int N> __device__ __forceinline__ uint32_t sign_word16(const uint32_t (&v)[N]) {   // one sign word from the 16 TMEM words v[OFF .. OFF + 15]
    constexpr uint32_t sel = ACC16 ? 0xFDB9u : 0xFBFBu;          // sign(byte 1, 3, 5, 7)  /  sign(byte 3, 7, 3, 7)
    uint32_t ea = prmt(v[OFF], v[OFF + 1], sel), eb = 0u;         // two independent chains
#pragma unroll
    for (int t = 1; t < 8; t += 2) {
        eb = imad(prmt(v[OFF + 2 * t], v[OFF + 2 * t + 1], sel), 1u << t, eb);
        if (t + 1 < 8) ea = imad(prmt(v[OFF + 2 * t + 2], v[OFF + 2 * t + 3], sel), 1u << (t + 1), ea);
    }
    return ea + eb;
}
template <bool ACC16> __device__ __forceinline__ void sign_words(const uint32_t (&v)[32], uint32_t& x0, uint32_t& x1) {
    x0 = sign_word16<ACC16, 0>(v); x1 = sign_word16<ACC16, 16>(v);
}
// Cheap pre-test of 16 TMEM words: the lane-wise maximum (VIMNMX3: three inputs per instruction, 8 instructions per 16 words, a
// quarter of the sign compaction's 32).  "Some accumulator of the 16 words has its sign bit clear" <=> some 16-bit half (ACC16:
// two S16 / FP16 accumulators per word) or the word itself (FP32 / S32) of the maximum is non-negative AS AN INTEGER -- a signed
// integer maximum is negative iff every input is, whatever the bits below the sign mean, so this is the same predicate as
// "sign word != kAllNegative" for all three accumulator types.
template <bool ACC16> __device__ __forceinline__ uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) {
    return ACC16 ? __vimax3_s16x2(a, b, c) : (uint32_t)__vimax3_s32((int)a, (int)b, (int)c);
}
template <bool ACC16, int OFF, int N> __device__ __forceinline__ uint32_t max16(const uint32_t (&v)[N]) {
    const uint32_t a = vmax3<ACC16>(v[OFF], v[OFF + 1], v[OFF + 2]), b = vmax3<ACC16>(v[OFF + 3], v[OFF + 4], v[OFF + 5]),
                   c = vmax3<ACC16>(v[OFF + 6], v[OFF + 7], v[OFF + 8]), d = vmax3<ACC16>(v[OFF + 9], v[OFF + 10], v[OFF + 11]),
                   e = vmax3<ACC16>(v[OFF + 12], v[OFF + 13], v[OFF + 14]);
    const uint32_t t1 = vmax3<ACC16>(a, b, c), t2 = vmax3<ACC16>(d, e, v[OFF + 15]);
    return vmax3<ACC16>(t1, t2, t2);
}
template <bool ACC16> __device__ __forceinline__ bool any_nonneg(uint32_t m) {
    return ACC16 ? ((m & 0x80008000u) != 0x80008000u) : ((int32_t)m >= 0);
}

// Per-epilogue-warp cursor into the raw-entry blocks.  next/left/blk are warp-uniform; `spare` (meaningful in lane 0) is
// the index of a block reserved AHEAD of time: the global atomic that reserves it is issued when the previous block is
// opened and its result is only consumed ~32 entries later, so its ~1 us latency never stalls the epilogue (a stalled
// epilogue warp stalls the whole MMA pipeline one tile later).
struct RawCursor { uint32_t* next; uint32_t left; uint32_t blk; uint32_t spare; };
// Open the pre-reserved block and reserve the one after it.  When the buffer is full the cursor points at the
// sacrificial block behind blk_cap, whose contents are never read: the host sees n_blocks > blk_cap and re-runs.
__device__ __forceinline__ void raw_new_block(RawCursor& rc, const TcParams& P, uint32_t lane, uint32_t tag) {      // (inlined: a call would spill the cursor to local memory)
    const uint32_t nb = __shfl_sync(0xffffffffu, rc.spare, 0);
    if (lane == 0) {
        if (rc.blk < P.blk_cap) P.blk_count[rc.blk] = kRawBlock - rc.left;          // close the old block
        if (nb < P.blk_cap) P.blk_tag[nb] = tag;                                    // a block holds entries of ONE column tile
        rc.spare = atomicAdd(P.n_blocks, 1u);
    }
    rc.blk = nb;
    rc.next = P.raw + (size_t)min(rc.blk, P.blk_cap) * (kRawBlock * kRawWords);
    rc.left = kRawBlock;
}
// Lanes whose window has a candidate in either of the warp's two chunks append one entry {window, column of chunk 0, its two
// sign masks, column of chunk 1, its two sign masks} to the warp's current block.  `t` is the ballot of those lanes (non-zero):
// one cursor update and one predicated, fire-and-forget 256-bit store (~2e-4 of the accumulators are candidates, so this
// runs for about every other tile of every warp).  One of the two chunks is almost always empty, but 16-byte entries of one chunk
// each were measured and rejected in round 2: choosing the chunk per lane and voting once more for the rare lane with both costs
// 0.4 ms of the 10.8 ms launch (two votes and two pushes: 1.0 ms) -- every instruction of this path competes with the issuing
// lane -- against the 0.3 ms the rescorer gains from reading half the bytes.
__device__ __forceinline__ void raw_push(RawCursor& rc, const TcParams& P, unsigned t, bool c, uint32_t win, uint32_t col0, uint32_t a0, uint32_t a1,
                                         uint32_t col1, uint32_t b0, uint32_t b1, uint32_t lane, uint32_t tag) {
    const uint32_t n = __popc(t);
    if (n > rc.left) raw_new_block(rc, P, lane, tag);       // n <= 32 <= kRawBlock
    uint32_t* d = rc.next + __popc(t & ((1u << lane) - 1u)) * kRawWords;
    if (!(TC_KNOCKOUT & 16))
        asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %8, 0;\n@q st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %7};\n}"
                     ::"l"(d), "r"(win), "r"(col0), "r"(a0), "r"(a1), "r"(col1), "r"(b0), "r"(b1), "r"((uint32_t)c) : "memory");
    rc.next += n * kRawWords;
    rc.left -= n;
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
// ZMASK = false: blocks of plain upper-case ACGT.  ZMASK = true: blocks with zero-contribution characters (lower case under the
// reference's BLAS-path semantics): their E entries are zeroed, the B image carries unshifted weights and one leading bias step
// (b200scan.cu: fold_z).  Each instance returns at once when the block is not of its kind.
// PAIR = true: launched as clusters of two CTAs (same TPC).  The pair takes an item of 2 * kTcSpan windows together -- rank 0 the
// first half of its window tiles, rank 1 the second -- and rank 0's issuer lane drives both tensor cores with cta_group::2 MMAs.
// I8 = true: INT8 operands (tcgen05.mma.kind::i8, K = 32: EIGHT motif positions per instruction instead of four).  An E entry
// then holds the one-hot bytes of four consecutive positions, E8[p] = [oh(p) oh(p+1) oh(p+2) oh(p+3)] with oh(c) = 1 << 8c, and
// "window r, K-chunk kk" is E8[g0 + r + 4 kk] (LBO = 64 B); the weights are integers  ceil(scale * (w - share))  clamped to
// [-127, 127] (b200scan.cu: fold_i8), the S32 accumulators are exact and stay inside +-2^15, so tcgen05.ld.pack::16b delivers
// their low halves as S16 and the FP16-accumulator epilogue (sign bits only) is reused unchanged (ACC16 = true).
template <bool ACC16, bool ZMASK, bool PAIR, bool I8 = false>
__global__ void __launch_bounds__(kTcThreads, kTcCtasPerSm)
filter_tc_kernel(TcParams P, BlockDev blk)
{
    static_assert(!I8 || (ACC16 && !PAIR), "INT8 operands: packed epilogue, single CTA");
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    constexpr uint32_t kBufs    = TC_BUFS;              // TMEM accumulator buffers
    constexpr uint32_t kBufCols = kTcMaxN;              // TMEM columns per buffer: one accumulator per column, FP32 or FP16
    constexpr uint32_t kColsPerWord = ACC16 ? 2 : 1;
    constexpr uint32_t kEpiPerQ = kTcEpiWarps / 4 / kTcEpiGroups;      // warps sharing one TMEM lane quarter of one tile

    extern __shared__ __align__(128) uint8_t smem_raw[];
    if ((__ldg(blk.has_zero) != 0) != ZMASK) return;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t*  sE     = smem;
    uint8_t*  sCodes = sE + kSmE;
    uint8_t*  sB     = sCodes + kSmCodes;
    uint64_t* sBars  = reinterpret_cast<uint64_t*>(sB + kSmB);
    // barrier map: [0..7] e_full, [8..15] e_empty, [16..19] t_full, [20..23] t_empty, [24] codes, [25] B
    volatile uint32_t* sMisc = reinterpret_cast<volatile uint32_t*>(sBars + 28);   // [0] work item, [1] tmem base
    uint4* sLut = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(sBars) + kSmBars);
    uint8_t* sZ = reinterpret_cast<uint8_t*>(sLut) + kSmLut;
    uint4* sOnes = reinterpret_cast<uint4*>(sZ + kSmZ);

    const uint32_t bars   = smem_u32(sBars);
    const uint32_t eFull  = bars, eEmpty = bars + 8 * 8, tFull = bars + 16 * 8, tEmpty = bars + 20 * 8;
    const uint32_t cBar   = bars + 24 * 8, bBar = bars + 25 * 8;
    // warp index through a lane-0 broadcast: tells ptxas the role branches below are warp-uniform, which is what
    // lets it use the uniform datapath (descriptors, barrier addresses) inside them
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // PAIR: eFull / tEmpty of the leader collect the arrivals of both CTAs; eEmpty / tFull are signalled in both by multicast commits
        for (uint32_t i = 0; i < kTcStages; i++) { mbar_init(eFull + 8 * i, (PAIR ? 2u : 1u) * kTcProducers); mbar_init(eEmpty + 8 * i, kTcMmaWarps); }
        for (uint32_t i = 0; i < kBufs; i++) { mbar_init(tFull + 8 * i, 1); mbar_init(tEmpty + 8 * i, (PAIR ? 2u : 1u) * kTcEpiWarps / kTcEpiGroups); }
        mbar_init(cBar, 1); mbar_init(bBar, 1);
        fence_mbar_init();
    }
    if (PAIR) cluster_sync();                    // both CTAs are resident and past their barrier initialisation
    if (warp == kTcProducers) { if (PAIR) tmem_alloc2(smem_u32(const_cast<uint32_t*>(&sMisc[1])), kBufs * kBufCols); else tmem_alloc(smem_u32(const_cast<uint32_t*>(&sMisc[1])), kBufs * kBufCols); }
    if (threadIdx.x < 16) {                      // one-hot FP16 (1.0 = 0x3C00) of code c in halves 0..3, of the next code in 4..7
        const uint32_t c0 = threadIdx.x & 3, c1 = threadIdx.x >> 2;
        const uint32_t v0 = 0x3C00u << (16 * (c0 & 1)), v1 = 0x3C00u << (16 * (c1 & 1));
        sLut[threadIdx.x] = make_uint4((c0 & 2) ? 0u : v0, (c0 & 2) ? v0 : 0u, (c1 & 2) ? 0u : v1, (c1 & 2) ? v1 : 0u);
    }
    if (ZMASK && threadIdx.x < kSmOnes / 16)                                                                // FP16: [1.0, 0, ..., 0] in every row; INT8: all ones
        sOnes[threadIdx.x] = I8 ? make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u) : make_uint4(0x00003C00u, 0u, 0u, 0u);
    if (ZMASK) fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = sMisc[1];
    // the leader's copies of the barriers that collect arrivals from both CTAs (for rank 0 these map to its own)
    const uint32_t eFullL = PAIR ? mapa_u32(eFull, 0) : eFull, tEmptyL = PAIR ? mapa_u32(tEmpty, 0) : tEmpty;
    (void)eFullL; (void)tEmptyL;

    long long ph_t = 0, ph_a[4] = {0, 0, 0, 0};   // B200_PHASE only (dead otherwise)
    (void)ph_t; (void)ph_a;
    uint32_t kE = 0;            // E stages produced / consumed so far (every role advances identically)
    uint32_t kT = 0;            // window tiles so far
    uint32_t nItem = 0, nBload = 0;
    int32_t  curTile = -1;
    RawCursor rawc; rawc.blk = 0xffffffffu; rawc.left = 0; rawc.next = P.raw; rawc.spare = 0;
    if (warp >= kTcEpiWarp0 && lane == 0) rawc.spare = atomicAdd(P.n_blocks, 1u);      // epilogue warps: reserve the first block early
    const uint32_t nItems = P.n_tiles * P.n_spans;
    // Epilogue role constants, made opaque to the optimiser: with 64 accumulator registers live it otherwise
    // re-derives them from threadIdx inside every tile iteration (each a ~5-cycle dependent ALU step for this warp).
    uint32_t eQ = warp & 3;                                            // TMEM lane quarter this warp may read
    uint32_t eGroup = ((warp - kTcEpiWarp0) >> 2) % kTcEpiGroups;      // which epilogue group (= TMEM buffer with 2 groups)
    uint32_t eWc0 = 32 * (((warp - kTcEpiWarp0) >> 2) / kTcEpiGroups); // first 32-word chunk of a tile this warp owns
    uint32_t eLaneAddr = tmem_base + ((32 * eQ) << 16);
    asm volatile("" : "+r"(eQ), "+r"(eGroup), "+r"(eWc0), "+r"(eLaneAddr));

    constexpr uint32_t kSpan = PAIR ? 2 * kTcSpan : kTcSpan;       // windows per item (P.n_spans counts these)
    while (true) {
        if (threadIdx.x == 0 && rank == 0) {
            const uint32_t it = atomicAdd(P.work_counter, 1u);
            sMisc[0] = it;
            if (PAIR) st_cluster_u32(mapa_u32(smem_u32(const_cast<uint32_t*>(&sMisc[0])), 1), it);     // the peer works on the same item
        }
        if (PAIR) cluster_sync(); else __syncthreads();
        const uint32_t item = sMisc[0];
        if (item >= nItems) break;
        const uint32_t t = item / P.n_spans, sp = item % P.n_spans;
        const TcTile tile = P.tiles[t];
        const uint32_t wItem = sp * kSpan;
        const uint32_t nTall = (min(kSpan, blk.n_payload - wItem) + 127) >> 7;          // window tiles of the item
        // PAIR: rank 0 takes the first ceil(nTall / 2) tiles, rank 1 the rest; both run nT iterations in lockstep (rank 1's last one
        // may be a phantom whose results are dropped)
        const uint32_t nT = PAIR ? (nTall + 1) >> 1 : nTall;
        const uint32_t myTiles = PAIR ? (rank ? nTall - nT : nT) : nTall;
        const uint32_t w0 = wItem + (rank ? 128 * nT : 0u);
        (void)myTiles;
        const uint32_t nSt = nT / kTcStageTiles + 1;            // E stages of the item: nT tiles + one halo tile
        const bool newTile = ((int32_t)t != curTile);
        curTile = (int32_t)t;

        if (threadIdx.x == 0) {
            // 2-bit codes of the span + one halo stage (+16 B so the last entry can see its successor)
            const uint32_t cbytes = (nT + 1) * 32 + 16, zbytes = ZMASK ? (nT + 1) * 16 + 16 : 0u;
            mbar_expect_tx(cBar, cbytes + zbytes);
            bulk_g2s(smem_u32(sCodes), reinterpret_cast<const uint8_t*>(blk.codes) + (w0 >> 2), cbytes, cBar);
            if (ZMASK) bulk_g2s(smem_u32(sZ), reinterpret_cast<const uint8_t*>(blk.zmask) + (w0 >> 3), zbytes, cBar);
            if (newTile) {                                    // PAIR: each CTA holds half of the tile's columns (a contiguous half of the image)
                const uint32_t bBytes = PAIR ? tile.b_bytes / 2 : tile.b_bytes, bOff = tile.b_off + rank * bBytes;
                mbar_expect_tx(bBar, bBytes);
                for (uint32_t off = 0; off < bBytes; off += 32768)
                    bulk_g2s(smem_u32(sB) + off, P.bimg + bOff + off, min(32768u, bBytes - off), bBar);
            }
        }

        if (warp < kTcProducers) {
            // ===================== producers: codes -> E ring (warp p fills entries 32p .. 32p+31 of every stage) =====================
            mbar_wait(cBar, nItem & 1, P.error_flag);
            if (PAIR && newTile) mbar_wait(bBar, nBload & 1, P.error_flag);      // PAIR: a stage reported full also vouches for this CTA's half of B
            const uint32_t nEnt = (nT + 1) * 128;                     // entries of the item incl. one halo tile
            for (uint32_t st = 0; st < nSt; st++) {
                const uint32_t k = kE + st, slot = k % kTcStages, ph = (k / kTcStages) & 1;
                PH_MARK();
                mbar_wait(eEmpty + 8 * slot, ph ^ 1, P.error_flag);
                PH_ACC(0);
                if (warp == 0) TC_TRACE(0, st, 0);
                if (!(TC_KNOCKOUT & 4) && !((TC_KNOCKOUT & 64) && t != 0 && nItem > 8)) {     // 64: fill only for column tile 0 (emulates an E-stationary loop order on random sequence; results invalid)
#pragma unroll
                    for (uint32_t r = 0; r < 4 * kTcStageTiles / kTcProducers; r++) {
                        const uint32_t e = 32 * (warp + r * kTcProducers) + lane;      // entry within the stage
                        const uint32_t g = st * kTcStageEnt + e;                       // entry within the item
                        if (g < nEnt) {
                            const uint32_t byte = g >> 2;
                            const uint32_t two = (uint32_t)sCodes[byte] | ((uint32_t)sCodes[byte + 1] << 8);
                            uint4 val;
                            if (I8) {                                                  // one-hot bytes of codes g .. g+3
                                const uint32_t c4 = two >> (2 * (g & 3));
                                val = make_uint4(1u << (8 * (c4 & 3u)), 1u << (8 * ((c4 >> 2) & 3u)), 1u << (8 * ((c4 >> 4) & 3u)), 1u << (8 * ((c4 >> 6) & 3u)));
                            } else val = sLut[(two >> (2 * (g & 3))) & 15u];           // [onehot(code g) | onehot(code g+1)]
                            if (ZMASK) {                                               // a masked character contributes an all-zero row
                                const uint32_t zz = ((uint32_t)sZ[g >> 3] | ((uint32_t)sZ[(g >> 3) + 1] << 8)) >> (g & 7);
                                if (I8) {
                                    if (zz & 1u) val.x = 0u;
                                    if (zz & 2u) val.y = 0u;
                                    if (zz & 4u) val.z = 0u;
                                    if (zz & 8u) val.w = 0u;
                                } else {
                                    if (zz & 1u) { val.x = 0u; val.y = 0u; }
                                    if (zz & 2u) { val.z = 0u; val.w = 0u; }
                                }
                            }
                            *reinterpret_cast<uint4*>(sE + (slot * kTcStageEnt + e) * 16) = val;
                            if (slot == 0 && e < kTcMirror)
                                *reinterpret_cast<uint4*>(sE + (kTcStages * kTcStageEnt + e) * 16) = val;
                        }
                    }
                }
                fence_proxy_async();              // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                // (PAIR: the stage lives in this CTA's shared memory and is read by this SM's tensor core; fence.proxy.async above made it
                //  visible to that proxy, so the arrival on the leader's barrier needs no cluster-wide release fence -- which costs ~600 cycles)
                if (lane == 0) { if (PAIR) mbar_arrive_cluster_cta(eFullL + 8 * slot); else mbar_arrive(eFull + 8 * slot); }
                PH_ACC(1); PH_COUNT();
                if (warp == 0) TC_TRACE(0, st, 1);
            }
        } else if (warp < kTcEpiWarp0) {
            // ===================== MMA issuer(s) =====================
            // This warp's instruction stream is on the critical path of every tile (alone on the SM: mbarrier.try_wait 43
            // cycles, tcgen05.commit 30, tcgen05.mma ~50 each -- tools/micro/issue_cost.cu -- and it shares its scheduler
            // with four epilogue warps), so it is kept minimal: one eFull wait and one eEmpty commit per STAGE of
            // kTcStageTiles tiles, one tEmpty wait and one tFull commit per tile, operands moved to uniform registers once
            // per tile.
            if (!PAIR && newTile) mbar_wait(bBar, nBload & 1, P.error_flag);
            const uint32_t n_k = __shfl_sync(0xffffffffu, tile.n_k, 0);
            // instruction descriptor: F16 x F16, D = F32 (c_format 1) or F16 (c_format 0), K-major A and B, M = 128 (256 for a CTA pair), N = n_pad
            // (INT8: signed 8-bit A and B -- a_format = b_format = 1 -- and D = S32, c_format 2)
            const uint32_t idesc = (I8 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (ACC16 ? 0u : (1u << 4))) | ((tile.n_pad >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
            const uint32_t nChunks = 2 * n_k;
            // the two 16-byte K-chunks of one MMA are 2 entries apart for FP16 (2 positions per entry), 4 for INT8 (4 positions per entry)
            constexpr uint32_t kChunkEnt = I8 ? 4u : 2u, kStepEnt = 2 * kChunkEnt;
            const uint64_t ad0 = umma_desc(smem_u32(sE), 16 * kChunkEnt, 128), bd0 = umma_desc(smem_u32(sB), 128, nChunks * 128);
            const uint32_t aLo0 = (uint32_t)ad0, aHi = (uint32_t)(ad0 >> 32), bLo0 = (uint32_t)bd0, bHi = (uint32_t)(bd0 >> 32);
            const uint32_t aLoOnes = (uint32_t)umma_desc(smem_u32(sOnes), 16 * kChunkEnt, 128);      // ZMASK: constant rows (the bias step)
            const uint32_t n_pos = ZMASK ? n_k - 1 : n_k;                                  // position steps (ZMASK: n_k counts the bias step too)
            (void)aLoOnes;
            // The whole tile loop runs in ONE elected lane, inside one branch: there ptxas moves the loop state to uniform
            // registers once per item and steps descriptors / barrier addresses with UIADD3 (predicating each tcgen05
            // instruction instead costs 5-7 R2UR moves in front of every one of them).
            if (kTcMmaWarps == 2) {
                // Two issuers, four or two TMEM buffers: issuer m takes the tiles with running index = m (mod 2).  A commit only tracks
                // the issuing thread's MMAs, so eEmpty counts two arrivals per stage: stage s is read by tiles 4s-1 .. 4s+3 (clipped
                // to the item); its last reader commits for its issuer, the reader before it for the other, a lone reader for both.
                const uint32_t mw = warp - kTcProducers;
                if (elect_one()) {
                    for (uint32_t i = (mw - kT) & 1u; i < nT; i += 2) {
                        const uint32_t st = i / kTcStageTiles, j = i % kTcStageTiles, k = kE + st, kt = kT + i, buf = kt % kBufs;
                        if (i < 2 || j < 2) mbar_wait(eFull + 8 * (k % kTcStages), (k / kTcStages) & 1, P.error_flag);       // this issuer's first tile in the stage
                        if (j + 1 == kTcStageTiles) mbar_wait(eFull + 8 * ((k + 1) % kTcStages), ((k + 1) / kTcStages) & 1, P.error_flag);
                        mbar_wait(tEmpty + 8 * buf, ((kt / kBufs) & 1) ^ 1, P.error_flag);
                        tc_fence_after();
                        const uint32_t d = tmem_base + buf * kBufCols;
                        uint32_t alo = aLo0 + (k % kTcStages) * kTcStageEnt + j * 128, blo = bLo0;
                        if (!(TC_KNOCKOUT & 2)) {
                            if (ZMASK) { umma_x<false, I8>(d, aLoOnes, aHi, blo, bHi, idesc, 0u); blo += 16; umma_x<false, I8>(d, alo, aHi, blo, bHi, idesc, 1u); }
                            else umma_x<false, I8>(d, alo, aHi, blo, bHi, idesc, 0u);
#pragma unroll 1
                            for (uint32_t m = 1; m < n_pos; m++) { alo += kStepEnt; blo += 16; umma_x<false, I8>(d, alo, aHi, blo, bHi, idesc, 1u); }
                        }
                        auto free_stage = [&](uint32_t s2) {
                            const uint32_t lo = s2 ? kTcStageTiles * s2 - 1 : 0u, hi = min(kTcStageTiles * s2 + kTcStageTiles - 1, nT - 1);
                            const uint32_t bar = eEmpty + 8 * ((kE + s2) % kTcStages);
                            if (i == hi) { umma_commit(bar); if (hi == lo) umma_commit(bar); }
                            else if (i + 1 == hi) umma_commit(bar);
                        };
                        free_stage(st);
                        if (j + 1 == kTcStageTiles && st + 1 < nSt) free_stage(st + 1);
                        umma_commit(tFull + 8 * buf);
                    }
                }
            } else
            if (TC_ISSUE_FAST && !PAIR && kBufs == 2 && kTcStageTiles == 4 && kTcStages == 4) {
                // Lean issue loop.  Measured with the B200_TRACE build (tools/tc_trace.py, round 2): the issuing lane never actually
                // waited for a TMEM buffer -- its own instruction stream (~70 SASS instructions per tile, each slowed by the four
                // epilogue warps on the same scheduler) set the tile period, while the epilogue groups idled half of the time.  So:
                // whole stages of four tiles from an unrolled body (tile-in-stage and buffer are static: A B A B), every MMA operand
                // except the E address loop-invariant, waits as inline spins (no CALL: uniform registers survive), parities flipped
                // in place, and nothing but the poll between a buffer's release and its next MMA.
                if (elect_one()) {
                    mbar_wait_spin(eFull + 8 * (kE % kTcStages), (kE / kTcStages) & 1, P.error_flag);
                    const uint32_t b0 = kT & 1u;                         // buffer of the item's first tile
                    const uint32_t dA = tmem_base + b0 * kBufCols, dB = tmem_base + (b0 ^ 1u) * kBufCols;
                    const uint32_t teA = tEmpty + 8 * b0, teB = tEmpty + 8 * (b0 ^ 1u), tfA = tFull + 8 * b0, tfB = tFull + 8 * (b0 ^ 1u);
                    uint32_t parA = ((kT >> 1) & 1u) ^ 1u, parB = (((kT + 1) >> 1) & 1u) ^ 1u;     // a buffer's parity flips with every use
                    // NP = position steps per tile as a compile-time constant (1..4: the MMA chain is straight-line code), 0 = run-time loop
                    auto run = [&](auto np_c) {
                        constexpr uint32_t NP = decltype(np_c)::value;
                        auto tileMma = [&](uint32_t d, uint32_t te, uint32_t par, uint32_t tf, uint32_t alo) {
                            mbar_wait_spin(te, par, P.error_flag);
                            tc_fence_after();
                            uint32_t blo = bLo0;
                            if (!(TC_KNOCKOUT & 2)) {
                                if (ZMASK) {
                                    umma_x<false, I8>(d, aLoOnes, aHi, blo, bHi, idesc, 0u);
                                    blo += 16;
                                    umma_x<false, I8>(d, alo, aHi, blo, bHi, idesc, 1u);
                                } else {
                                    umma_x<false, I8>(d, alo, aHi, blo, bHi, idesc, 0u);
                                }
                                if (NP) {
#pragma unroll
                                    for (uint32_t m = 1; m < NP; m++) umma_x<false, I8>(d, alo + m * kStepEnt, aHi, blo + 16 * m, bHi, idesc, 1u);
                                } else {
#pragma unroll 1
                                    for (uint32_t m = 1; m < n_pos; m++) {
                                        alo += kStepEnt; blo += 16;
                                        umma_x<false, I8>(d, alo, aHi, blo, bHi, idesc, 1u);
                                    }
                                }
                            }
                            umma_commit(tf);
                        };
                        uint32_t kCur = kE;
#pragma unroll 1
                        for (uint32_t st = nT >> 2; st != 0; st--) {
                            const uint32_t slot = kCur % kTcStages, nslot = (kCur + 1) % kTcStages;
                            const uint32_t aS = aLo0 + slot * kTcStageEnt;                    // address fields are in 16-byte units = entries
                            tileMma(dA, teA, parA, tfA, aS);        parA ^= 1u;
                            tileMma(dB, teB, parB, tfB, aS + 128);  parB ^= 1u;
                            tileMma(dA, teA, parA, tfA, aS + 256);  parA ^= 1u;
                            mbar_wait_spin(eFull + 8 * nslot, ((kCur + 1) / kTcStages) & 1, P.error_flag);      // the last tile's halo lies in the next stage
                            tileMma(dB, teB, parB, tfB, aS + 384);  parB ^= 1u;
                            // MMAs complete in issue order: with the stage's last tile done, so is every reader of the stage
                            umma_commit(eEmpty + 8 * slot);
                            if (st == 1 && (nT & 3u) == 0) umma_commit(eEmpty + 8 * nslot);    // the item's last tile also frees its halo stage
                            kCur++;
                        }
                        const uint32_t rem = nT & 3u;                        // the last span of a block: a partial stage (tiles + halo all inside it)
                        if (rem) {
                            const uint32_t slot = kCur % kTcStages;
                            const uint32_t aS = aLo0 + slot * kTcStageEnt;
                            tileMma(dA, teA, parA, tfA, aS);
                            if (rem > 1) tileMma(dB, teB, parB, tfB, aS + 128);
                            if (rem > 2) tileMma(dA, teA, parA ^ 1u, tfA, aS + 256);
                            umma_commit(eEmpty + 8 * slot);
                        }
                    };
                    switch (n_pos) {
                        case 1: run(std::integral_constant<uint32_t, 1>{}); break;
                        case 2: run(std::integral_constant<uint32_t, 2>{}); break;
                        case 3: run(std::integral_constant<uint32_t, 3>{}); break;
                        case 4: run(std::integral_constant<uint32_t, 4>{}); break;
                        default: run(std::integral_constant<uint32_t, 0>{}); break;
                    }
                }
            } else
            if (rank == 0 && elect_one()) {           // PAIR: the leader's lane issues for both CTAs
                mbar_wait(eFull + 8 * (kE % kTcStages), (kE / kTcStages) & 1, P.error_flag);
                // loop state is stepped incrementally (adds and masks only: every instruction of this lane is on the critical path)
                uint32_t j = 0, kCur = kE, kt = kT;                      // tile within the stage, stage counter, tile counter
                uint32_t aOff = (kE % kTcStages) * kTcStageEnt;          // first E entry of the tile, in ring entries
#pragma unroll 1
                for (uint32_t left = nT; left != 0; left--) {
                    const uint32_t buf = kt % kBufs;
                    const bool lastOfStage = (j + 1 == kTcStageTiles), lastTile = (left == 1);
                    // everything the MMAs need is computed BEFORE the wait for the TMEM buffer: what follows the wake-up is on the
                    // critical path of the hand-back chain (release -> issue -> MMA -> tFull -> load -> release)
                    const uint32_t d = tmem_base + buf * kBufCols;
                    // A: window rows straight out of E (Toeplitz): row r, chunk kk -> E[entry0 + r + 2*kk]; one MMA = 4 entries = 64 B
                    // B: [8-column group][chunk] blocks of 128 B: one MMA = 2 chunks = 256 B
                    uint32_t alo = aLo0 + aOff, blo = bLo0;               // address fields are in 16-byte units = entries
                    const uint32_t teBar = tEmpty + 8 * buf, tePar = ((kt / kBufs) & 1) ^ 1, tfBar = tFull + 8 * buf;
                    PH_MARK();
                    if (lastOfStage)                                      // its halo lies in the next stage
                        mbar_wait(eFull + 8 * ((kCur + 1) % kTcStages), ((kCur + 1) / kTcStages) & 1, P.error_flag);
                    TC_TRACE(1, nT - left, 0);
                    PH_ACC(0);
                    mbar_wait(teBar, tePar, P.error_flag);
                    PH_ACC(1);
                    TC_TRACE(1, nT - left, 1);
                    tc_fence_after();
                    if (!(TC_KNOCKOUT & 2)) {
                        if (ZMASK) {                                      // step 0: D = bias (constant one-hot rows x chunk 0 of B)
                            umma_x<PAIR, I8>(d, aLoOnes, aHi, blo, bHi, idesc, 0u);
                            blo += 16;
                            umma_x<PAIR, I8>(d, alo, aHi, blo, bHi, idesc, 1u);
                        } else {
                            umma_x<PAIR, I8>(d, alo, aHi, blo, bHi, idesc, 0u);
                        }
#pragma unroll 1
                        for (uint32_t m = 1; m < n_pos; m++) {
                            alo += kStepEnt; blo += 16;
                            umma_x<PAIR, I8>(d, alo, aHi, blo, bHi, idesc, 1u);
                        }
                    }
                    // tFull first: the epilogue group waiting for it is the critical path.  MMAs complete in issue order: once the last
                    // tile of a stage is done, so is every reader of the stage (its own tiles and the halo read of the stage before).
                    // The item's last tile also frees its halo stage.
                    commit_x<PAIR>(tfBar);
                    if (lastOfStage || lastTile) commit_x<PAIR>(eEmpty + 8 * (kCur % kTcStages));
                    if (lastOfStage && lastTile) commit_x<PAIR>(eEmpty + 8 * ((kCur + 1) % kTcStages));
                    PH_ACC(2); PH_COUNT();
                    TC_TRACE(1, nT - left, 2);
                    aOff = (aOff + 128) % (kTcStages * kTcStageEnt);
                    kt++;
                    if (lastOfStage) { j = 0; kCur++; } else j++;
                }
            }
            __syncwarp();
        } else {
            // ===================== epilogue: TMEM -> sign test -> candidates =====================
            const uint32_t nWords = tile.n_pad / kColsPerWord;        // 32-bit TMEM columns of a tile (multiple of 32)
            // tag of the raw blocks this warp fills: the whole tile, or -- in the specialised loop below, where a warp owns one 128-column
            // half of every tile -- that half (the fused rescorer then keeps only 128 columns' weights in shared memory)
            const bool halfTags = TC_HALF_TAGS && TC_EPI_FAST && ACC16 && !PAIR && kTcEpiGroups == 2 && kBufs == 2 && kEpiPerQ == 2 && !(TC_KNOCKOUT & 1) && nWords == 128;
            const uint32_t halfCol0 = 4 * eWc0, halfCols = tile.n_cols > halfCol0 ? min(128u, tile.n_cols - halfCol0) : 0u;
            const uint32_t tileTag = halfTags ? (((tile.col0 + halfCol0) << 9) | halfCols) : ((tile.col0 << 9) | tile.n_cols);
            if (newTile && rawc.blk != 0xffffffffu) {                 // another column tile: close the open block (a block holds ONE tile's entries)
                if (lane == 0 && rawc.blk < P.blk_cap) P.blk_count[rawc.blk] = kRawBlock - rawc.left;
                rawc.blk = 0xffffffffu; rawc.left = 0;
            }
            // with two groups, group g takes the tiles whose running index is g mod 2 (= the tiles of TMEM buffer g)
            const uint32_t i0 = (kTcEpiGroups > 1) ? ((eGroup - kT) & (kTcEpiGroups - 1)) : 0u;
            uint32_t kt = kT + i0, win0 = w0 + 128 * i0 + 32 * eQ;    // running tile index; window of lane 0
            constexpr bool kFastOk = TC_EPI_FAST && ACC16 && !PAIR && kTcEpiGroups == 2 && kBufs == 2 && kEpiPerQ == 2 && !(TC_KNOCKOUT & 1);
            if (kFastOk && nWords == 128) {
                // ---- full 256-column tile, packed accumulators: this warp owns 64 consecutive words (128 columns) of every tile of ITS
                // buffer (group g <-> buffer g: the tiles with running index = g mod 2), so TMEM address and barriers are loop constants
                // and nothing but the barrier poll stands between tFull and the load.
                const uint32_t buf = kt & 1u;
                const uint32_t fullBar = tFull + 8 * buf, emptyBar = tEmpty + 8 * buf;
                const uint32_t taddr = eLaneAddr + buf * kBufCols + 4 * eWc0;           // eWc0 = 32 * sub: words 64 sub .. 64 sub + 63 = TMEM columns 128 sub ..
                const uint32_t colA = tile.col0 + 4 * eWc0, colB = colA + 64;
                uint32_t tph = (kt >> 1) & 1u, win = win0 + lane;
                for (uint32_t i = i0; i < nT; i += 2, tph ^= 1u, win += 256) {
                    PH_MARK();
                    mbar_wait_spin(fullBar, tph, P.error_flag);
                    PH_ACC(0);
                    if (warp == kTcEpiWarp0) TC_TRACE(2, i, 0); else if (warp == kTcEpiWarp0 + 7) TC_TRACE(3, i, 0);
                    tc_fence_after();
                    uint32_t v[64];
                    tmem_ld64_pack16(taddr, v);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(emptyBar);
                    PH_ACC(1);
                    if (warp == kTcEpiWarp0) TC_TRACE(2, i, 2); else if (warp == kTcEpiWarp0 + 7) TC_TRACE(3, i, 2);
                    const uint32_t g0 = max16<true, 0>(v), g1 = max16<true, 16>(v), g2 = max16<true, 32>(v), g3 = max16<true, 48>(v);
                    const uint32_t gm = vmax3<true>(g0, g1, vmax3<true>(g2, g3, g3));
                    const bool c = win < blk.n_payload && any_nonneg<true>(gm);
                    const unsigned t = (TC_KNOCKOUT & 8) ? 0u : __ballot_sync(0xffffffffu, c);
                    if (t) {                                               // about every other pass: compact only the groups that hold a candidate
                        uint32_t a0 = kAllNegative, a1 = kAllNegative, b0 = kAllNegative, b1 = kAllNegative;
                        if (__any_sync(0xffffffffu, any_nonneg<true>(g0))) a0 = sign_word16<true, 0>(v);
                        if (__any_sync(0xffffffffu, any_nonneg<true>(g1))) a1 = sign_word16<true, 16>(v);
                        if (__any_sync(0xffffffffu, any_nonneg<true>(g2))) b0 = sign_word16<true, 32>(v);
                        if (__any_sync(0xffffffffu, any_nonneg<true>(g3))) b1 = sign_word16<true, 48>(v);
                        raw_push(rawc, P, t, c, win, colA, a0, a1, colB, b0, b1, lane, tileTag);
                    }
                    PH_ACC(2); PH_COUNT();
                    if (warp == kTcEpiWarp0) TC_TRACE(2, i, 3); else if (warp == kTcEpiWarp0 + 7) TC_TRACE(3, i, 3);
                }
            } else
            for (uint32_t i = i0; i < nT; i += kTcEpiGroups, kt += kTcEpiGroups, win0 += 128 * kTcEpiGroups) {
                const uint32_t buf = kt % kBufs, tph = (kt / kBufs) & 1;
                PH_MARK();
                mbar_wait(tFull + 8 * buf, tph, P.error_flag);
                PH_ACC(0);
                if (warp == kTcEpiWarp0) TC_TRACE(2, i, 0); else if (warp == kTcEpiWarp0 + 7) TC_TRACE(3, i, 0);
                tc_fence_after();
                const bool winOk = (!PAIR || i < myTiles) && win0 + lane < blk.n_payload;
                const uint32_t taddr = eLaneAddr + buf * kBufCols;
                // this warp owns the 32-word chunks sub, sub + kEpiPerQ, sub + 2 kEpiPerQ, ... of the tile, taken two at a
                // time (both loads in flight).  As soon as its LAST loads have landed in registers the warp hands the
                // TMEM buffer back, before looking at the data.
                constexpr uint32_t step = 32 * kEpiPerQ;
                // (instances that also carry the fast loop above take ONE chunk per iteration here: two 32-register loads next to the
                // fast loop's 64-register one make ptxas spill; this loop then only sees narrow tiles)
                constexpr uint32_t kPerIter = kFastOk ? 1u : 2u;
                bool released = false;
                for (uint32_t wc = eWc0; wc < ((TC_KNOCKOUT & 1) ? 0u : nWords); wc += kPerIter * step) {
                    const bool has1 = kPerIter == 2 && wc + step < nWords;
                    uint32_t v0[32], v1[32];
                    tmem_ld_words<ACC16>(taddr, wc, v0);
                    if (kPerIter == 2 && has1) tmem_ld_words<ACC16>(taddr, wc + step, v1);
                    tmem_ld_wait();
                    if (wc + kPerIter * step >= nWords) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR) mbar_arrive_cluster_cta(tEmptyL + 8 * buf); else mbar_arrive(tEmpty + 8 * buf); }
                        PH_ACC(1);
                        released = true;
                        if (warp == kTcEpiWarp0) TC_TRACE(2, i, 2); else if (warp == kTcEpiWarp0 + 7) TC_TRACE(3, i, 2);
                    }
                    // both compactions first (independent chains interleave in the issue slots), then ONE test and ONE vote
                    uint32_t a0, a1, b0 = kAllNegative, b1 = kAllNegative;
                    sign_words<ACC16>(v0, a0, a1);
                    if (kPerIter == 2 && has1) sign_words<ACC16>(v1, b0, b1);
                    const bool c = winOk && (((a0 ^ kAllNegative) | (a1 ^ kAllNegative) | (b0 ^ kAllNegative) | (b1 ^ kAllNegative)) != 0u);
                    const unsigned t = (TC_KNOCKOUT & 8) ? 0u : __ballot_sync(0xffffffffu, c);
                    if (t) raw_push(rawc, P, t, c, win0 + lane, (tile.col0 + wc * kColsPerWord) | (ACC16 ? 0u : kRawFp32Flag), a0, a1,
                                    (tile.col0 + (wc + step) * kColsPerWord) | (ACC16 ? 0u : kRawFp32Flag), b0, b1, lane, tileTag);
                }
                if (!released) {                                      // this warp owns no chunk of such a narrow tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_cluster_cta(tEmptyL + 8 * buf); else mbar_arrive(tEmpty + 8 * buf); }
                }
                PH_ACC(2); PH_COUNT();
                if (warp == kTcEpiWarp0) TC_TRACE(2, i, 3); else if (warp == kTcEpiWarp0 + 7) TC_TRACE(3, i, 3);
            }
        }
        kE += nSt;
        kT += nT;
        nItem++;
        if (newTile) nBload++;
        if (PAIR) cluster_sync(); else __syncthreads();          // item boundary: every role (of both CTAs) is done with sCodes / sB, the pipelines are drained
    }

#ifdef B200_PHASE
    if (lane == 0 && (warp == 0 || warp == kTcProducers || warp == kTcEpiWarp0)) {
        const uint32_t role = warp == 0 ? 0u : (warp == kTcProducers ? 1u : 2u);
        for (int k = 0; k < 4; k++) atomicAdd(P.trace + 16 * role + k, (unsigned long long)ph_a[k]);
    }
#endif
    if (warp >= kTcEpiWarp0 && lane == 0) {                                          // close the open block and the unused spare
        if (rawc.blk < P.blk_cap) P.blk_count[rawc.blk] = kRawBlock - rawc.left;
        if (rawc.spare < P.blk_cap) P.blk_count[rawc.spare] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();
    if (warp == kTcProducers) { if (PAIR) tmem_dealloc2(tmem_base, kBufs * kBufCols); else tmem_dealloc(tmem_base, kBufs * kBufCols); }
}

} // namespace b200
