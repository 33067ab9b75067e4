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
//    in shared memory; a CTA sweeps windows, D[128 windows x N columns] accumulates in TMEM (FP32), double
//    buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//  * The tensor pass is a CONSERVATIVE FILTER: weights are  fp16_round_up(W[j][b] - thr'/L)  with
//    thr' = thr - margin, so  acc >= 0  whenever the exact FP32 score >= thr (b200scan.cu: build_tc_tiles).
//    The epilogue only tests sign bits (AND-reduction, half a LOP3 per score) and appends (pos, col)
//    candidates; rescore.cuh then recomputes those few scores exactly.  R never exists.
//
// Roofline: tensor pipe.  One tcgen05.mma (M=128, N, K=16) covers 4 motif positions of N columns for 128
// windows and takes N/2 cycles; the epilogue must drain 128 x N FP32 accumulators per tile from TMEM.
//
// Warp roles (320 threads, 1 CTA/SM, persistent with an atomic work counter):
//   warp 0      producer: codes -> E ring (8 stages of 128 entries + mirrored halo)
//   warp 1      TMEM allocation, B/codes bulk loads are issued by thread 0; lane 0 issues tcgen05.mma
//   warps 2..9  epilogue, two warps per TMEM lane quarter (they alternate 32-column chunks so that each SM
//               sub-partition always has an independent instruction stream to issue from): tcgen05.ld
//               32x32b.x32 (two in flight), tree-shaped AND of the sign bits, candidate staging + flush
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int      kTcThreads  = 320;
constexpr uint32_t kTcEpiWarps = 8;
constexpr uint32_t kTcSpan     = 32768;      // windows per work item
constexpr uint32_t kTcStages   = 8;          // E ring stages (128 entries = 2 KB each)
constexpr uint32_t kTcMirror   = 64;         // entries mirrored past the ring end (>= 2*(2*nK_max-1))
constexpr uint32_t kTcMaxN     = 256;
constexpr uint32_t kTcStageCap = 64;         // staged candidates per epilogue warp

struct TcTile {
    uint32_t col0;      // first sorted column
    uint32_t n_cols;    // real columns
    uint32_t n_pad;     // N of the MMA, multiple of 32, <= 256 (padding columns can never pass the filter)
    uint32_t n_k;       // MMAs per 128-window tile = ceil(Lmax / 4)
    uint32_t b_off;     // byte offset of the tile's B image in TcParams::bimg
    uint32_t b_bytes;   // n_pad * (2*n_k) * 16
};

struct TcParams {
    const uint8_t* bimg;
    const TcTile*  tiles;
    uint32_t       n_tiles;
    uint32_t       n_spans;
    unsigned int*  work_counter;
    Cand*          cand;
    unsigned long long* n_cand;
    unsigned long long  cand_cap;
    unsigned int*  error_flag;
};

// shared memory carve-up (bytes)
constexpr uint32_t kSmE      = (kTcStages * 128 + kTcMirror) * 16;            // 17408
constexpr uint32_t kSmCodes  = kTcSpan / 4 + 128;                              //  8320
constexpr uint32_t kSmB      = kTcMaxN * (2 * (kMaxLen / 4)) * 16;             // 131072
constexpr uint32_t kSmStage  = kTcEpiWarps * kTcStageCap * 8;                  //  4096
constexpr uint32_t kSmBars   = 32 * 8;
constexpr uint32_t kTcSmemBytes = kSmE + kSmCodes + kSmB + kSmStage + kSmBars + 128;

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle ("interleave") shared memory descriptor: rows 16 B apart inside an 8-row core matrix,
// 8-row groups SBO apart, the two 16-byte K-chunks of one K=16 MMA LBO apart (cute mma_traits_sm100.hpp,
// canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; version = 1 on sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}


// AND of 32 accumulators as a depth-4 tree of 3-input LOP3s: sign bit set <=> all 32 are negative.
__device__ __forceinline__ uint32_t and_tree32(const uint32_t (&v)[32]) {
    const uint32_t t0 = v[0] & v[1] & v[2],    t1 = v[3] & v[4] & v[5],    t2 = v[6] & v[7] & v[8],    t3 = v[9] & v[10] & v[11];
    const uint32_t t4 = v[12] & v[13] & v[14], t5 = v[15] & v[16] & v[17], t6 = v[18] & v[19] & v[20], t7 = v[21] & v[22] & v[23];
    const uint32_t t8 = v[24] & v[25] & v[26], t9 = v[27] & v[28] & v[29], t10 = v[30] & v[31];
    const uint32_t u0 = t0 & t1 & t2, u1 = t3 & t4 & t5, u2 = t6 & t7 & t8, u3 = t9 & t10;
    return (u0 & u1) & (u2 & u3);
}
// bit (31 - j) = sign bit of v[j]; four independent funnel-shift chains of 8
__device__ __forceinline__ uint32_t sign_mask32(const uint32_t (&v)[32]) {
    uint32_t m[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int c = 0; c < 4; c++) m[c] = __funnelshift_l(v[8 * c + j], m[c], 1);
    }
    return (m[0] << 24) | (m[1] << 16) | (m[2] << 8) | m[3];
}

struct CandStage {          // per-warp candidate staging (warp-uniform count in a register)
    Cand* buf; uint32_t n;
};
__device__ __forceinline__ void stage_flush(CandStage& st, const TcParams& P, uint32_t lane) {
    __syncwarp();
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.n_cand, (unsigned long long)st.n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t s = lane; s < st.n; s += 32)
        if (base + s < P.cand_cap) P.cand[base + s] = st.buf[s];
    __syncwarp();
    st.n = 0;
}
// One 32-column chunk of this thread's window: fast sign test, rare slow path that lists the candidates.
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], bool winOk, uint32_t win, uint32_t col_base,
                                               CandStage& st, const TcParams& P, uint32_t lane) {
    const uint32_t a = and_tree32(v);
    if (!__any_sync(0xffffffffu, winOk && (int32_t)a >= 0)) return;
    uint32_t c = winOk ? ~sign_mask32(v) : 0u;          // bit (31-j) set <=> accumulator j >= 0 <=> candidate
    while (true) {                                      // rounds: every lane with candidates left pushes one (ascending column)
        const bool has = c != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (!bal) break;
        if (has) {
            const int b = 31 - __clz(c);
            c &= ~(1u << b);
            Cand cd; cd.pos = win; cd.col = col_base + (31 - b);
            st.buf[st.n + __popc(bal & ((1u << lane) - 1u))] = cd;
        }
        st.n += __popc(bal);
        if (st.n > kTcStageCap - 32) stage_flush(st, P, lane);      // one global atomic per >= 32 candidates
    }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
filter_tc_kernel(TcParams P, BlockDev blk)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    if (__ldg(blk.has_zero) != 0) return;                       // zero-mask blocks take the gather kernel

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t*  sE     = smem;
    uint8_t*  sCodes = sE + kSmE;
    uint8_t*  sB     = sCodes + kSmCodes;
    Cand*     sStage = reinterpret_cast<Cand*>(sB + kSmB);
    uint64_t* sBars  = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStage) + kSmStage);
    // barrier map: [0..7] e_full, [8..15] e_empty, [16,17] t_full, [18,19] t_empty, [20] codes, [21] B
    volatile uint32_t* sMisc = reinterpret_cast<volatile uint32_t*>(sBars + 24);   // [0] work item, [1] tmem base

    const uint32_t bars   = smem_u32(sBars);
    const uint32_t eFull  = bars, eEmpty = bars + 8 * 8, tFull = bars + 16 * 8, tEmpty = bars + 18 * 8;
    const uint32_t cBar   = bars + 20 * 8, bBar = bars + 21 * 8;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < kTcStages; i++) { mbar_init(eFull + 8 * i, 1); mbar_init(eEmpty + 8 * i, 1); }
        for (uint32_t i = 0; i < 2; i++) { mbar_init(tFull + 8 * i, 1); mbar_init(tEmpty + 8 * i, kTcEpiWarps); }
        mbar_init(cBar, 1); mbar_init(bBar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(&sMisc[1])), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sMisc[1];

    uint32_t kE = 0;            // E stages produced / consumed so far (every role advances identically)
    uint32_t kT = 0;            // window tiles so far
    uint32_t nItem = 0, nBload = 0;
    int32_t  curTile = -1;
    CandStage stage;            // epilogue: staged candidates of this warp (count is warp-uniform)
    stage.buf = sStage + ((warp >= 2) ? (warp - 2) : 0) * kTcStageCap; stage.n = 0;
    const uint32_t nItems = P.n_tiles * P.n_spans;

    while (true) {
        if (threadIdx.x == 0) sMisc[0] = atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const uint32_t item = sMisc[0];
        if (item >= nItems) break;
        const uint32_t t = item / P.n_spans, sp = item % P.n_spans;
        const TcTile tile = P.tiles[t];
        const uint32_t w0 = sp * kTcSpan;
        const uint32_t nwin = min(kTcSpan, blk.n_payload - w0);
        const uint32_t nT = (nwin + 127) >> 7;
        const bool newTile = ((int32_t)t != curTile);
        curTile = (int32_t)t;

        if (threadIdx.x == 0) {
            // 2-bit codes of the span + one halo stage (+16 B so the last entry can see its successor)
            const uint32_t cbytes = (nT + 1) * 32 + 16;
            mbar_expect_tx(cBar, cbytes);
            bulk_g2s(smem_u32(sCodes), reinterpret_cast<const uint8_t*>(blk.codes) + (w0 >> 2), cbytes, cBar);
            if (newTile) {
                mbar_expect_tx(bBar, tile.b_bytes);
                for (uint32_t off = 0; off < tile.b_bytes; off += 32768)
                    bulk_g2s(smem_u32(sB) + off, P.bimg + tile.b_off + off, min(32768u, tile.b_bytes - off), bBar);
            }
        }

        if (warp == 0) {
            // ===================== producer: codes -> E ring =====================
            mbar_wait(cBar, nItem & 1, P.error_flag);
            for (uint32_t i = 0; i <= nT; i++) {
                const uint32_t k = kE + i, slot = k % kTcStages, ph = (k / kTcStages) & 1;
                mbar_wait(eEmpty + 8 * slot, ph ^ 1, P.error_flag);
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
                    const uint32_t e = 32 * q + lane;                 // entry within the stage
                    const uint32_t byte = 32 * i + (e >> 2);
                    const uint32_t two = (uint32_t)sCodes[byte] | ((uint32_t)sCodes[byte + 1] << 8);
                    const uint32_t c0 = (two >> (2 * (e & 3))) & 3u, c1 = (two >> (2 * (e & 3) + 2)) & 3u;
                    const uint32_t v0 = 0x3C00u << (16 * (c0 & 1)), v1 = 0x3C00u << (16 * (c1 & 1));
                    uint4 val;
                    val.x = (c0 & 2) ? 0u : v0;  val.y = (c0 & 2) ? v0 : 0u;
                    val.z = (c1 & 2) ? 0u : v1;  val.w = (c1 & 2) ? v1 : 0u;
                    *reinterpret_cast<uint4*>(sE + (slot * 128 + e) * 16) = val;
                    if (slot == 0 && e < kTcMirror)
                        *reinterpret_cast<uint4*>(sE + (kTcStages * 128 + e) * 16) = val;
                }
                fence_proxy_async();              // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(eFull + 8 * slot);
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (newTile) mbar_wait(bBar, nBload & 1, P.error_flag);
            const uint32_t idesc = (1u << 4) | ((tile.n_pad >> 3) << 17) | ((128u >> 4) << 24);   // F16 x F16 -> F32, K-major A and B
            const uint32_t nChunks = 2 * tile.n_k;
            const uint32_t bBase = smem_u32(sB), eBase = smem_u32(sE);
            for (uint32_t i = 0; i < nT; i++) {
                const uint32_t k = kE + i, slot = k % kTcStages, ph = (k / kTcStages) & 1;
                const uint32_t k1 = k + 1, slot1 = k1 % kTcStages, ph1 = (k1 / kTcStages) & 1;
                const uint32_t kt = kT + i, buf = kt & 1, tph = (kt >> 1) & 1;
                if (i == 0) mbar_wait(eFull + 8 * slot, ph, P.error_flag);     // later tiles waited for it as their halo stage
                mbar_wait(eFull + 8 * slot1, ph1, P.error_flag);
                mbar_wait(tEmpty + 8 * buf, tph ^ 1, P.error_flag);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t d = tmem_base + buf * kTcMaxN;
                    for (uint32_t m = 0; m < tile.n_k; m++) {
                        // A: window rows straight out of E (Toeplitz): row r, chunk kk -> E[slot*128 + r + 2*kk]
                        const uint64_t ad = umma_desc(eBase + (slot * 128 + 4 * m) * 16, 32, 128);
                        // B: [8-column group][chunk] blocks of 128 B: chunks 128 B apart, column groups nChunks*128 B apart
                        const uint64_t bd = umma_desc(bBase + (2 * m) * 128, 128, nChunks * 128);
                        umma_f16(d, ad, bd, idesc, m > 0 ? 1u : 0u);
                    }
                    umma_commit(eEmpty + 8 * slot);       // stage k is dead once tile i (and i-1 before it) completed
                    umma_commit(tFull + 8 * buf);
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit(eEmpty + 8 * ((kE + nT) % kTcStages));   // the halo stage
            __syncwarp();
        } else {
            // ===================== epilogue: TMEM -> sign test -> candidates =====================
            const uint32_t q = warp & 3;                              // TMEM lane quarter this warp may read
            const uint32_t half = (warp - 2) >> 2;                    // which of the quarter's two warps
            for (uint32_t i = 0; i < nT; i++) {
                const uint32_t kt = kT + i, buf = kt & 1, tph = (kt >> 1) & 1;
                mbar_wait(tFull + 8 * buf, tph, P.error_flag);
                tc_fence_after();
                const uint32_t win = w0 + 128 * i + 32 * q + lane;
                const bool winOk = win < blk.n_payload;
                const uint32_t taddr = tmem_base + ((32 * q) << 16) + buf * kTcMaxN;
                uint32_t cc = 32 * half;                              // this warp takes chunks half, half+2, half+4, ...
                for (; cc + 64 < tile.n_pad; cc += 128) {             // two chunks per trip, both loads in flight
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + cc, v0);
                    tmem_ld32(taddr + cc + 64, v1);
                    tmem_ld_wait();
                    epilogue_chunk(v0, winOk, win, tile.col0 + cc, stage, P, lane);
                    epilogue_chunk(v1, winOk, win, tile.col0 + cc + 64, stage, P, lane);
                }
                if (cc < tile.n_pad) {
                    uint32_t v0[32];
                    tmem_ld32(taddr + cc, v0);
                    tmem_ld_wait();
                    epilogue_chunk(v0, winOk, win, tile.col0 + cc, stage, P, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tEmpty + 8 * buf);
            }
        }
        kE += nT + 1;
        kT += nT;
        nItem++;
        if (newTile) nBload++;
        __syncthreads();          // item boundary: every role is done with sCodes / sB / the pipelines are drained
    }

    if (warp >= 2 && stage.n) stage_flush(stage, P, lane);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

} // namespace b200
