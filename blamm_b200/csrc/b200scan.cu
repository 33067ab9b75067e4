// libb200scan.so -- context management, motif packing, kernel launches and the C ABI of include/b200scan.h.
// Everything device-side is sm_100a CUDA in the .cuh files next to this one; there is no CPU scoring path.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>            // header-only NVTX 3: ranges show up in nsys / ncu timelines, cost nothing without a tool attached
#define B200_NVTX 1
#endif
#endif
#include "common.cuh"
#include "filter_tc.cuh"
#include "gather.cuh"
#include "order.cuh"
#include "pack.cuh"
#include "rescore.cuh"

using namespace b200;

namespace {

thread_local std::string g_create_error;

// NVTX range for the lifetime of the object: the host-side phases of the ABI (submit = upload + launches, collect = wait + hit
// download, set_motifs = folding + tile planning) in a profiler's timeline (tracing)
struct NvtxRange {
#ifdef B200_NVTX
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char*) {}
#endif
};

constexpr uint32_t kPadBytes = 16384;             // slack behind every device sequence buffer (window / span over-reads)
constexpr size_t   kGatherSmemW = 64 * 1024;      // FP32 weights per gather column tile

struct Slot {
    // host staging (pinned).  Everything but the counters is allocated at the slot's first use (a caller that alternates two of
    // the three slots pays for two; the ASCII staging buffers exist only for callers that submit characters).
    uint8_t*  h_ascii = nullptr;                  // pageable b200scan_submit_ascii sources are staged here
    uint32_t* h_frag = nullptr;  size_t frag_cap = 0;
    void*     h_hits = nullptr;  size_t h_hit_bytes = 0;      // sized from the blocks actually collected (grows by 1/4 steps)
    uint32_t* h_bucket = nullptr;  size_t h_bucket_cap = 0;   // B200SCAN_HITS_8: host copy of bucket_start
    unsigned long long* h_counters = nullptr;     // [0] n_cand, [1] n_hits, [2] has_zero|error (as 2 x u32)
    // device
    uint8_t*  d_ascii = nullptr;
    uint32_t* d_codes = nullptr;
    uint32_t* d_zmask = nullptr;
    uint32_t* d_frag = nullptr;
    b200scan_hit* d_hits = nullptr;  unsigned long long hit_cap = 0;
    uint32_t* d_bucket_start = nullptr;  size_t bucket_cap = 0;      // B200SCAN_HITS_8: bucket_start of the block in flight
    unsigned long long* d_counters = nullptr;     // [0] n_cand, [1] n_hits, [2] {has_zero, error_flag}, [3] {work_counter, -}
    // state
    bool in_flight = false, resident = false;
    uint64_t n_total = 0, n_payload = 0, n_frag = 0;
    bool packed_zero_known = false;               // submit_packed: has_zero decided on the host
    int  hit_bytes = B200SCAN_HITS_16;            // record format of the block in flight / resident (b200scan_set_hit_format)
    cudaEvent_t ev[10] = {};                      // upload stream: 0 start, 1 after h2d, 2 after pack; compute stream: 8 scoring starts,
                                                  // 3 after score, 4 after rescore, 9 after ordering, 5 after counters d2h; copy stream: 6/7 hit d2h
    b200scan_timing timing = {};
};

} // namespace

struct b200scan_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // hit downloads: overlap the next block's kernels
    cudaStream_t up_stream = nullptr;        // block uploads + packing: overlap the previous block's kernels
    uint64_t max_block = 0;
    unsigned long long max_hits = 0;         // initial per-slot hit capacity (b200scan_create)
    unsigned long long hit_budget = 0;       // largest hit / candidate capacity a regrow may reach (b200scan_create: from the free device memory)
    double margin16_scale = 1.0;             // debug knob (env B200SCAN_MARGIN16_SCALE): scales the FP16-accumulation error bound
    double i8_max_overshoot = 16.0;          // INT8 operands for a tile iff every column's WORST-CASE overshoot (L / scale, score units) stays below (env B200SCAN_I8_MAX_OVERSHOOT; 0 = never).  Measured on the bench set the mean overshoot is far below the bound: 1.20 candidates per hit with every tile on INT8, against 1.44 for FP16 accumulators
    int engine = B200SCAN_ENGINE_AUTO;
    int hit_format = B200SCAN_HITS_16;
    std::string err;
    Slot slot[B200SCAN_NUM_SLOTS];
    // candidates are shared by the slots (stream order serialises filter -> rescore per block)
    unsigned long long cand_cap = 0;         // raw entries the filter may write (sizes d_raw: blk_cap blocks of kRawBlock entries)
    // raw entries of the tensor filter (blocks of kRawBlock entries of kRawWords words), also shared by the slots
    uint32_t* d_raw = nullptr;  uint32_t* d_blk_count = nullptr;  uint32_t* d_blk_tag = nullptr;  uint32_t blk_cap = 0;
    uint32_t fuse_max_w = 0;      // shared-memory weight positions of the fused kernel for the loaded motif set
    int fuse_ctas_per_sm = 1;     // its CTAs that fit an SM with that much shared memory
    // motifs
    bool have_motifs = false;
    uint32_t n_cols = 0, max_len = 0;  uint64_t sum_len = 0;
    float4* d_w = nullptr;  uint32_t *d_woff = nullptr, *d_len = nullptr, *d_orig = nullptr;  float* d_thr = nullptr;
    GatherTile* d_gtiles = nullptr;  std::vector<GatherTile> gtiles;  size_t gather_smem = 0;
    TcTile* d_ttiles = nullptr;  std::vector<TcTile> ttiles;  uint8_t* d_bimg = nullptr;
    uint32_t n_tiles8 = 0;        // ttiles[0 .. n_tiles8) use INT8 operands,
    uint32_t n_tiles16 = 0;       // the next n_tiles16 FP16 operands with FP16 accumulators, the rest FP32 accumulators
    TcTile* d_ttiles_z = nullptr;  uint8_t* d_bimg_z = nullptr;  uint32_t n_tiles16_z = 0, n_tiles8_z = 0;     // the same for blocks with zero-contribution characters
    bool tc_usable = false;
    int  tc_acc_bits = 32;        // accumulators of the tensor tiles: 16, 32, or 0 when tiles differ
    bool pair_mode = false;       // filter launched as CTA pairs (cta_group::2 MMAs)
    int  acc_pref = 0;            // 0 auto, 8 (INT8 operands), 16, 32 (b200scan_set_tensor_accumulator)
    double cand_inflation = 0;    // mean margin over columns (diagnostic)
    uint8_t* d_flush = nullptr;
    unsigned long long* d_trace = nullptr;   // B200_TRACE builds only
    // B200SCAN_HITS_8 (order.cuh): per-bucket counters / cursors and the bucket-contiguous scratch list, shared by the slots
    // (the ordering kernels of a block run back to back on the compute stream)
    uint32_t *d_bucket_cnt = nullptr, *d_bucket_cursor = nullptr, *d_coarse_start = nullptr;  size_t bucket_cap = 0;     // per coarse bucket
    Hit12* d_sort_tmp = nullptr;  unsigned long long sort_cap = 0;
    // empirical histograms
    uint32_t hist_bins = 0;  unsigned long long* d_hist = nullptr;  float *d_hmin = nullptr, *d_hwid = nullptr;
    GatherTile* d_htiles = nullptr;  std::vector<GatherTile> htiles;  size_t hist_smem = 0;
    int hist_kernel = 3;          // 3: gather_hist2_kernel (position-row weights, lanes count different columns at a time); 2: the same with
                                  // match-aggregated counts; 1 / 0: gather_hist_kernel with / without aggregated counts (env
                                  // B200SCAN_HIST_KERNEL: comparison and independent check)
    HistTile2* d_htiles2 = nullptr;  std::vector<HistTile2> htiles2;
    std::vector<uint32_t> h_len_sorted, h_woff_sorted, h_orig_sorted;      // host copies of the sorted column metadata
};

namespace {

int fail(b200scan_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do { cudaError_t e_ = (call);                                                                        \
         if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? B200SCAN_ENOMEM : B200SCAN_ECUDA, \
                                            "%s -> %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

template <class T> void dfree(T*& p) { if (p) cudaFree(p); p = nullptr; }
template <class T> void hfree(T*& p) { if (p) cudaFreeHost(p); p = nullptr; }

// smallest FP16 value >= x (x finite), as raw bits
uint16_t half_round_up(double x)
{
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    __half h = __float2half_rn(f);
    uint16_t bits; std::memcpy(&bits, &h, 2);
    if (__half2float(h) < f) {
        if (bits == 0x8000u) bits = 0x0001u;            // -0 -> smallest positive
        else if (bits & 0x8000u) bits -= 1;             // negative: towards zero
        else bits += 1;                                 // positive: away from zero
    }
    return bits;
}

// ---------------------------------------------------------------------------------------------------------
// Conservative folding of ONE motif column into tensor-core filter weights (host code; no device needed).
// build_motifs calls these per column; b200scan_debug_fold exposes them to the CPU tests, which check the recall
// guarantee  "in-order FP32 score >= threshold  =>  accumulator >= 0"  on exhaustive and boundary windows.
// ---------------------------------------------------------------------------------------------------------
// Conservative folding.  With x_j the FP32 weights picked by a window (exact in-order FP32 score s = sum x_j + e1,
// |e1| <= (L-1) 2^-24 A, A = sum_j max|x_j|), the tensor accumulator is  acc = sum y_j + e2  with
//     y_j = fp16_up(x_j - thr'/L) >= x_j - thr'/L,      thr' = thr - margin,
// so  s >= thr  =>  acc >= margin - |e1| - |e2|, and the filter is exact-recall when margin >= |e1| + |e2|.
//   FP32 accumulators: |e2| <= L 2^-18 (A'+1)  (generous for any FP32-ish accumulate, A' = sum_j max|y_j|).
//   FP16 accumulators: D is rounded to FP16 after every MMA (4 positions).  For a window that is a true hit
//     (final sum >= 0) the partial sum after step k lies in [-R_k, M_k] (R_k = largest possible remaining sum,
//     M_k = largest possible prefix), so |D_k| <= B_k = max(M_k, R_k, 0).  Allowing a rounding of one FP16 ulp
//     (2^-10 relative: covers round-to-nearest and truncation) on EVERY internal add of the 4 products and the
//     old accumulator:  |e2| <= 2^-10 * 4 * sum_k (B_{k-1} + A_k),  A_k = sum of max|y_j| over the step.
struct Folded { std::vector<uint16_t> y; double margin; bool always; uint16_t bias = 0;
                std::vector<int8_t> q; int32_t qbias = 0; bool never = false; };      // INT8 operands (fold_i8)
Folded fold_col(const float4* wc, uint32_t L, float thr, bool acc16, double margin16_scale)
{
    Folded f; f.y.assign(4 * L, 0); f.margin = 0; f.always = false;
    double A = 0;
    bool finite = std::isfinite(thr);
    for (uint32_t j = 0; j < L; j++) {
        const float4 v = wc[j];
        const float a4[4] = {v.x, v.y, v.z, v.w};
        double m = 0;
        for (float a : a4) { if (!std::isfinite(a)) finite = false; m = std::max(m, std::fabs((double)a)); }
        A += m;
    }
    if (!finite) { f.always = true; return f; }
    const double e1 = (L - 1) * std::ldexp(A, -24);
    double margin = 1e-3 + e1 + L * std::ldexp(A + std::fabs((double)thr) + 1.0, -18);
    for (int iter = 0; iter < 6; iter++) {
        const double share = ((double)thr - margin) / L;
        std::vector<double> ymax(L), yabs(L);
        for (uint32_t j = 0; j < L; j++) {
            const float4 v = wc[j];
            const float a4[4] = {v.x, v.y, v.z, v.w};
            double mx = -1e300, ma = 0;
            for (uint32_t o = 0; o < 4; o++) {
                const double y = (double)a4[o] - share;
                if (std::fabs(y) > 30000.0) { f.always = true; return f; }
                const uint16_t h = half_round_up(y);
                f.y[4 * j + o] = h;
                __half hh; std::memcpy(&hh, &h, 2);
                const double yr = (double)__half2float(hh);
                mx = std::max(mx, yr); ma = std::max(ma, std::fabs(yr));
            }
            ymax[j] = mx; yabs[j] = ma;
        }
        if (!acc16) { f.margin = margin; return f; }
        // FP16 accumulation bound
        const uint32_t nk = (L + 3) / 4;
        std::vector<double> pre(nk + 1, 0.0), suf(nk + 1, 0.0);
        for (uint32_t k = 1; k <= nk; k++) { pre[k] = pre[k - 1]; for (uint32_t j = 4 * (k - 1); j < std::min(L, 4 * k); j++) pre[k] += ymax[j]; }
        for (uint32_t k = nk; k-- > 0;) { suf[k] = suf[k + 1]; for (uint32_t j = 4 * k; j < std::min(L, 4 * (k + 1)); j++) suf[k] += ymax[j]; }
        double sum = 0, tot_abs = 0;
        for (uint32_t k = 1; k <= nk; k++) {
            double Ak = 0; for (uint32_t j = 4 * (k - 1); j < std::min(L, 4 * k); j++) Ak += yabs[j];
            const double Bprev = (k == 1) ? 0.0 : std::max(std::max(pre[k - 1], suf[k - 1]), 0.0);
            sum += Bprev + Ak; tot_abs += Ak;
        }
        if (tot_abs > 30000.0) { f.always = true; return f; }
        const double need = 1e-3 + e1 + margin16_scale * std::ldexp(4.0 * sum, -10);
        if (need <= margin) { f.margin = margin; return f; }
        margin = need * 1.02;
    }
    f.margin = 1e9;           // did not converge: the caller falls back to FP32 accumulators
    return f;
}

// Blocks with zero-contribution characters (lower case under the reference's BLAS-path semantics, sequence.cpp:312-319)
// use a second image: the weights are NOT shifted by a threshold share (a masked position must add exactly 0); instead
// one extra leading MMA step adds the bias  b = fp16_up(-(thr - margin))  to every window through a constant one-hot
// operand.  acc = b + sum y_j >= score - thr + margin as before, for either sign of the threshold, and a fully masked
// window gets acc = b < 0 instead of a spurious candidate.  FP16 accumulation: D_0 = b exactly; for a true hit the
// partial sum after position step k lies in [-R_k, b + M_k] with M_k / R_k built from max(y, 0) (a masked position
// contributes 0), so |D_k| <= max(R_k, |b| + M_k) and the same per-add ulp bound applies.
Folded fold_col_z(const float4* wc, uint32_t L, float thr, bool acc16, double margin16_scale)
{
    Folded f; f.y.assign(4 * L, 0); f.margin = 0; f.always = false; f.bias = 0;
    double A = 0;
    bool finite = std::isfinite(thr);
    std::vector<double> ymax(L), ymin(L), yabs(L);
    for (uint32_t j = 0; j < L; j++) {
        const float4 v = wc[j];
        const float a4[4] = {v.x, v.y, v.z, v.w};
        double m = 0, mx = 0, mn = 0, ma = 0;              // mx / mn start at 0: the masked contribution
        for (uint32_t o = 0; o < 4; o++) {
            if (!std::isfinite(a4[o])) { finite = false; continue; }
            m = std::max(m, std::fabs((double)a4[o]));
            if (std::fabs((double)a4[o]) > 30000.0) { finite = false; continue; }
            const uint16_t h = half_round_up((double)a4[o]);
            f.y[4 * j + o] = h;
            __half hh; std::memcpy(&hh, &h, 2);
            const double yr = (double)__half2float(hh);
            mx = std::max(mx, yr); mn = std::min(mn, yr); ma = std::max(ma, std::fabs(yr));
        }
        A += m; ymax[j] = mx; ymin[j] = mn; yabs[j] = ma;
    }
    if (!finite || std::fabs((double)thr) > 30000.0) { f.always = true; return f; }
    const double e1 = (L - 1) * std::ldexp(A, -24);
    double margin = 1e-3 + e1 + L * std::ldexp(A + std::fabs((double)thr) + 1.0, -18);
    const uint32_t nk = (L + 3) / 4;
    // prefix maxima / minima and suffix maxima of the position sums at the step boundaries
    std::vector<double> pre(nk + 1, 0.0), pmin(nk + 1, 0.0), suf(nk + 1, 0.0);
    for (uint32_t k = 1; k <= nk; k++) {
        pre[k] = pre[k - 1]; pmin[k] = pmin[k - 1];
        for (uint32_t j = 4 * (k - 1); j < std::min(L, 4 * k); j++) { pre[k] += ymax[j]; pmin[k] += ymin[j]; }
    }
    for (uint32_t k = nk; k-- > 0;) { suf[k] = suf[k + 1]; for (uint32_t j = 4 * k; j < std::min(L, 4 * (k + 1)); j++) suf[k] += ymax[j]; }
    for (int iter = 0; iter < 6; iter++) {
        f.bias = half_round_up(-((double)thr - margin));
        f.margin = margin;
        if (!acc16) return f;
        __half hb; std::memcpy(&hb, &f.bias, 2);
        const double b = (double)__half2float(hb);
        double sum = 0;
        for (uint32_t k = 1; k <= nk; k++) {
            double Ak = 0; for (uint32_t j = 4 * (k - 1); j < std::min(L, 4 * k); j++) Ak += yabs[j];
            // D_{k-1} lies in [max(-R, b + Pmin), b + Pmax] for a window that ends up >= 0
            const double hi = b + pre[k - 1], lo = std::max(-suf[k - 1], b + pmin[k - 1]);
            sum += std::max(std::fabs(hi), std::fabs(lo)) + Ak;
        }
        const double need = 1e-3 + e1 + margin16_scale * std::ldexp(4.0 * sum, -10);
        if (need <= margin) return f;
        margin = need * 1.02;
    }
    f.margin = 1e9;           // did not converge: FP32 accumulators for this tile
    return f;
}

// INT8 operands (filter_tc_kernel<.., I8 = true>: eight positions per MMA, exact S32 accumulation).  With x_j the FP32
// weights a window picks, S their real sum and s the reference's in-order FP32 sum (|s - S| <= e1), a hit s >= thr has
// S >= thr' = thr - e1 - 1e-3.  The integer weights are  q_j[b] = max(-127, ceil(scale * (x_j[b] - share_j)))  with
// sum_j share_j = thr':  acc = sum q_j >= scale * (S - thr') >= 0 -- rounding up and clamping up can only ADD candidates,
// so no accumulation margin is needed at all; what it costs is an overshoot of < 1 / scale per position.
//   share_j = best_j - c, c = slack / L, slack = sum_j best_j - thr'  (every position's best letter gets the same value c),
//   scale = 127 / max(c, slack - c + 1/8):  a letter that loses more than the whole slack kills a window on its own; it may
//   clamp at -127, everything milder is represented to 1 / scale.  |acc| <= 127 L < 2^15.
// zmode (blocks with zero-contribution characters): unshifted weights q = ceil(scale * x), the bias step adds
//   B = max(-4064, ceil(-scale * thr')) spread over the 32 K rows of one MMA; a masked position contributes exactly 0;
//   slack is taken over max(best_j, 0).  Returns the worst-case overshoot L / scale (score units) as `margin`.
Folded fold_col_i8(const float4* wc, uint32_t L, float thr, int zmode)
{
    Folded f; f.q.assign(4 * L, 0); f.margin = 0; f.always = false;
    double A = 0, smax = 0, pmax = 0;
    bool finite = std::isfinite(thr);
    std::vector<double> best(L);
    for (uint32_t j = 0; j < L; j++) {
        const float4 v = wc[j];
        const float a4[4] = {v.x, v.y, v.z, v.w};
        double m = 0, b = -1e300;
        for (float a : a4) { if (!std::isfinite(a)) finite = false; m = std::max(m, std::fabs((double)a)); b = std::max(b, (double)a); }
        A += m; best[j] = zmode ? std::max(b, 0.0) : b; smax += best[j]; pmax = std::max(pmax, b);
    }
    if (!finite || A > 1e6 || std::fabs((double)thr) > 1e6) { f.always = true; return f; }
    const double thrp = (double)thr - (L - 1) * std::ldexp(A, -24) - 1e-3;
    const double slack = smax - thrp;
    if (slack < 0) { f.never = true; return f; }                   // no window can reach the threshold
    const double c = slack / L;
    const double scale = zmode ? 127.0 / std::max(std::max(pmax, 0.0) + 1e-9, slack + 0.125)
                               : 127.0 / std::max(c + 1e-9, slack - c + 0.125);
    if (!(scale > 1e-3) || L / scale > 64.0) { f.always = true; f.margin = 1e9; return f; }      // threshold far below the best score: INT8 cannot resolve it (auto mode then keeps FP16 operands)
    for (uint32_t j = 0; j < L; j++) {
        const float4 v = wc[j];
        const float a4[4] = {v.x, v.y, v.z, v.w};
        const double share = zmode ? 0.0 : best[j] - c;
        for (uint32_t o = 0; o < 4; o++) {
            const double q = std::ceil(scale * ((double)a4[o] - share));
            f.q[4 * j + o] = (int8_t)std::min(127.0, std::max(-127.0, q));      // q <= 127 by the choice of scale (min: guards the last ulp)
            if (q > 127.0) { f.always = true; f.margin = 1e9; return f; }
        }
    }
    if (zmode) {
        const double b = std::ceil(-scale * thrp);
        if (b > 4064.0) { f.always = true; f.margin = 1e9; return f; }
        f.qbias = (int32_t)std::max(-4064.0, b);
    }
    f.margin = L / scale;
    return f;
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core tiling: sorted columns are cut into tiles of <= 256 columns (N padded to 32) minimising
//   sum over tiles of  max(MMA cycles, epilogue cycles)  per 128-window tile.  This replaces the reference's
// greedy zero-area tiling of P (MotifContainer::generateMatrixTiles, motif.cpp:482-540) -- like there, the
// split is a pure performance device and cannot change results.
// ---------------------------------------------------------------------------------------------------------
void plan_tc_tiles(const std::vector<uint32_t>& len_sorted, bool acc16, uint32_t pos_per_mma, std::vector<std::pair<uint32_t, uint32_t>>& out)
{
    const uint32_t n = (uint32_t)len_sorted.size();
    // epilogue cycles per column per 128-window tile (ALU-pipe bound, from ncu: profiles/), MMA = n_k * N / 2
    const double kEpiPerCol = acc16 ? 1.1 : 2.9, kFixed = 96.0;
    std::vector<double> best(n + 1, 1e300);
    std::vector<uint32_t> from(n + 1, 0);
    best[0] = 0;
    for (uint32_t j = 1; j <= n; j++) {
        const uint32_t nk = (len_sorted[j - 1] + pos_per_mma - 1) / pos_per_mma;
        for (uint32_t i = (j > kTcMaxN ? j - kTcMaxN : 0); i < j; i++) {
            const uint32_t npad = ((j - i) + 63) / 64 * 64;
            const double c = best[i] + std::max(nk * npad * 0.5, kEpiPerCol * npad) + kFixed;
            if (c < best[j]) { best[j] = c; from[j] = i; }
        }
    }
    out.clear();
    for (uint32_t j = n; j > 0; j = from[j]) out.push_back({from[j], j});
    std::reverse(out.begin(), out.end());
}

int build_motifs(b200scan_ctx* ctx, const float* P, int32_t ldp, int32_t n_cols, const int32_t* col_len, const float* thr)
{
    // stable sort by length
    std::vector<uint32_t> order(n_cols);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return col_len[a] < col_len[b]; });

    std::vector<uint32_t> len(n_cols), woff(n_cols), orig(n_cols);
    std::vector<float> thr_s(n_cols);
    uint64_t sum_len = 0; uint32_t max_len = 0;
    for (int32_t s = 0; s < n_cols; s++) {
        const uint32_t c = order[s];
        len[s] = (uint32_t)col_len[c]; woff[s] = (uint32_t)sum_len; orig[s] = c; thr_s[s] = thr[c];
        sum_len += len[s]; max_len = std::max(max_len, len[s]);
    }
    std::vector<float4> w(sum_len);
    for (int32_t s = 0; s < n_cols; s++) {
        const float* col = P + (size_t)orig[s] * ldp;
        for (uint32_t j = 0; j < len[s]; j++)
            w[woff[s] + j] = make_float4(col[4 * j], col[4 * j + 1], col[4 * j + 2], col[4 * j + 3]);
    }

    // ---- gather tiles: as many columns as fit the shared-memory weight budget ----
    std::vector<GatherTile> gt;
    size_t max_meta = 0;
    for (uint32_t s = 0; s < (uint32_t)n_cols;) {
        GatherTile t{s, 0, woff[s], 0};
        while (s < (uint32_t)n_cols && (size_t)(t.n_w + len[s]) * 16 <= kGatherSmemW && t.n_cols < 1024) {
            t.n_w += len[s]; t.n_cols++; s++;
        }
        gt.push_back(t);
        max_meta = std::max(max_meta, (size_t)t.n_cols * 12);
    }
    ctx->gather_smem = kGatherSmemW + max_meta;

    // ---- tensor-core tiles and the FP16 B image ----  (the per-column folding rules: fold_col / fold_col_z / fold_col_i8 above)
    auto fold    = [&](uint32_t sc, bool acc16) { return fold_col(&w[woff[sc]], len[sc], thr_s[sc], acc16, ctx->margin16_scale); };
    auto fold_z  = [&](uint32_t sc, bool acc16) { return fold_col_z(&w[woff[sc]], len[sc], thr_s[sc], acc16, ctx->margin16_scale); };
    auto fold_i8 = [&](uint32_t sc, int zmode) { return fold_col_i8(&w[woff[sc]], len[sc], thr_s[sc], zmode); };

    // Accumulator type PER TILE: FP16 accumulators (half the epilogue work) where every column of the tile keeps its margin
    // <= 2 score units, FP32 otherwise -- a few long or extreme motifs then cost their own tile, not the whole set.
    // Operand type PER TILE as well: INT8 operands (half the MMAs) where the worst-case overshoot of every column of the tile
    // stays <= ctx->i8_max_overshoot score units (the bound is pessimistic, see its definition; it guards the exact rescorer
    // against columns whose threshold lies far below their best score); the FP16-operand kinds otherwise.  acc_pref 8 forces
    // INT8 everywhere, 16 / 32 exclude it.
    const bool try16 = ctx->acc_pref != 32 && max_len <= (uint32_t)kMaxLen;
    const bool try8 = (ctx->acc_pref == 0 || ctx->acc_pref == 8) && max_len <= (uint32_t)kMaxLen && ctx->i8_max_overshoot > 0;
    std::vector<std::pair<uint32_t, uint32_t>> cuts;
    plan_tc_tiles(len, try16, try8 ? 8u : 4u, cuts);
    bool tc_ok = true;
    double margin_sum = 0;
    // zmode 0: blocks without zero-contribution characters; zmode 1: with (bias step first, see fold_z)
    auto build_image = [&](int zmode, std::vector<TcTile>& tt, std::vector<uint8_t>& bimg, uint32_t& n16, uint32_t& n8) {
        std::vector<Folded> folded(n_cols);                    // filled tile by tile: the FP16 folds (3/4 of the host time of a set_motifs when
                                                               // computed for every column) only for tiles that do not end up on INT8 operands
        std::vector<uint32_t> tile_acc16(cuts.size(), 0);      // 1 FP16 accumulators, 0 FP32, 2 INT8 operands
        for (size_t ti = 0; ti < cuts.size(); ti++) {
            if (try8) {
                std::vector<Folded> f8;
                bool ok8 = true;
                for (uint32_t sc = cuts[ti].first; sc < cuts[ti].second && ok8; sc++) {
                    f8.push_back(fold_i8(sc, zmode));
                    if (ctx->acc_pref != 8 && !f8.back().never && f8.back().margin > ctx->i8_max_overshoot) ok8 = false;     // (a non-finite column has margin 0: 'always' under every kind)
                }
                if (ok8) {
                    for (uint32_t sc = cuts[ti].first; sc < cuts[ti].second; sc++) folded[sc] = std::move(f8[sc - cuts[ti].first]);
                    tile_acc16[ti] = 2u;
                    continue;
                }
            }
            for (uint32_t sc = cuts[ti].first; sc < cuts[ti].second; sc++) folded[sc] = zmode ? fold_z(sc, try16) : fold(sc, try16);
            bool ok = try16;
            if (ok && ctx->acc_pref != 16)
                for (uint32_t sc = cuts[ti].first; sc < cuts[ti].second; sc++)
                    if (!folded[sc].always && folded[sc].margin > 2.0) { ok = false; break; }
            tile_acc16[ti] = ok ? 1u : 0u;
            for (uint32_t sc = cuts[ti].first; sc < cuts[ti].second; sc++) {
                if (!ok && try16) folded[sc] = zmode ? fold_z(sc, false) : fold(sc, false);
                // forced FP16 accumulators: a column whose bound did not converge goes to the exact rescorer wholesale
                if (ok && !folded[sc].always && folded[sc].margin > 1e8) folded[sc].always = true;
            }
        }
        if (!zmode) for (const auto& f : folded) if (!f.always && !f.never && f.margin < 1e8) margin_sum += f.margin;
        n16 = 0; n8 = 0;
        for (uint32_t a : tile_acc16) { n16 += (a == 1u); n8 += (a == 2u); }
        tt.clear(); bimg.clear();
        for (size_t ti = 0; ti < cuts.size(); ti++) {
            const auto& cut = cuts[ti];
            TcTile t{};
            t.acc16 = tile_acc16[ti];
            t.col0 = cut.first; t.n_cols = cut.second - cut.first;
            t.n_pad = (t.n_cols + 63) / 64 * 64;
            const uint32_t ppm = t.acc16 == 2u ? 8u : 4u;                    // positions per MMA
            t.n_k = (len[cut.second - 1] + ppm - 1) / ppm + (zmode ? 1u : 0u);      // MMA steps per window tile (zmode: + the bias step)
            const uint32_t nChunks = 2 * t.n_k;
            t.b_off = (uint32_t)bimg.size();
            t.b_bytes = t.n_pad * nChunks * 16;
            bimg.resize(bimg.size() + t.b_bytes, 0);
            if (t.acc16 == 2u) {
                // INT8 image: [8-column group][chunk][8 columns][16 bytes = 4 positions x ACGT]; zmode: chunks 0-1 = the bias step
                int8_t* img8 = reinterpret_cast<int8_t*>(bimg.data() + t.b_off);
                auto at8 = [&](uint32_t n, uint32_t kk, uint32_t byte) -> int8_t& { return img8[(((n >> 3) * nChunks + kk) * 8 + (n & 7)) * 16 + byte]; };
                for (uint32_t n = 0; n < t.n_pad; n++) {
                    const Folded* f = n < t.n_cols ? &folded[t.col0 + n] : nullptr;
                    if (!f || f->never || f->always) {                   // padding / unreachable column: acc < 0 for every window; degenerate column: acc > 0
                        const int8_t v = (f && f->always) ? 1 : -127;
                        if (zmode) at8(n, 0, 0) = v; else for (uint32_t o = 0; o < 4; o++) at8(n, 0, o) = v;
                        continue;
                    }
                    if (zmode) {                                         // bias spread over the 32 K rows of the bias step
                        const int32_t b = f->qbias, each = b / 32, rem = b - 32 * each;       // |each| <= 127; rem has the sign of b
                        for (uint32_t r = 0; r < 32; r++)
                            at8(n, r >> 4, r & 15) = (int8_t)(each + ((int32_t)r < std::abs(rem) ? (rem > 0 ? 1 : -1) : 0));
                    }
                    for (uint32_t j = 0; j < len[t.col0 + n]; j++)
                        for (uint32_t o = 0; o < 4; o++) at8(n, (j >> 2) + (zmode ? 2u : 0u), (j & 3) * 4 + o) = f->q[4 * j + o];
                }
                tt.push_back(t);
                continue;
            }
            uint16_t* img = reinterpret_cast<uint16_t*>(bimg.data() + t.b_off);
            auto at = [&](uint32_t n, uint32_t j, uint32_t o) -> uint16_t& {        // column n (tile-local), position j, letter o
                const uint32_t kk = (j >> 1) + (zmode ? 2u : 0u);
                return img[(((n >> 3) * nChunks + kk) * 8 + (n & 7)) * 8 + (j & 1) * 4 + o];
            };
            auto bias = [&](uint32_t n) -> uint16_t& { return img[(((n >> 3) * nChunks) * 8 + (n & 7)) * 8]; };   // zmode: chunk 0, half 0
            for (uint32_t n = 0; n < t.n_pad; n++) {
                if (n >= t.n_cols) {                         // padding column: can never reach acc >= 0
                    if (zmode) bias(n) = 0xFBFFu; else for (uint32_t o = 0; o < 4; o++) at(n, 0, o) = 0xFBFFu;   // -65504
                    continue;
                }
                const Folded& f = folded[t.col0 + n];
                if (f.always) {                              // degenerate column: every window goes to the exact rescorer
                    if (zmode) bias(n) = 0x3C00u; else for (uint32_t o = 0; o < 4; o++) at(n, 0, o) = 0x3C00u;   // acc = +1
                    continue;
                }
                if (zmode) bias(n) = f.bias;
                for (uint32_t j = 0; j < len[t.col0 + n]; j++)
                    for (uint32_t o = 0; o < 4; o++) at(n, j, o) = f.y[4 * j + o];
            }
            tt.push_back(t);
        }
        // launch order: INT8 tiles, FP16-accumulator tiles, FP32-accumulator tiles
        std::stable_sort(tt.begin(), tt.end(), [](const TcTile& a, const TcTile& b) {
            auto rank = [](uint32_t k) { return k == 2u ? 0 : (k == 1u ? 1 : 2); };
            return rank(a.acc16) < rank(b.acc16); });
    };
    std::vector<TcTile> tt, tt_z;
    std::vector<uint8_t> bimg, bimg_z;
    uint32_t n16 = 0, n16_z = 0, n8 = 0, n8_z = 0;
    build_image(0, tt, bimg, n16, n8);
    build_image(1, tt_z, bimg_z, n16_z, n8_z);
    ctx->n_tiles16 = n16; ctx->n_tiles16_z = n16_z; ctx->n_tiles8 = n8; ctx->n_tiles8_z = n8_z;
    if (max_len > (uint32_t)kMaxLen) tc_ok = false;
    ctx->tc_acc_bits = (n8 == cuts.size()) ? 8 : (n16 == cuts.size()) ? 16 : (n16 == 0 && n8 == 0 ? 32 : 0);
    ctx->cand_inflation = n_cols ? margin_sum / n_cols : 0;

    // ---- upload ----
    dfree(ctx->d_w); dfree(ctx->d_woff); dfree(ctx->d_len); dfree(ctx->d_orig); dfree(ctx->d_thr);
    dfree(ctx->d_gtiles); dfree(ctx->d_ttiles); dfree(ctx->d_bimg); dfree(ctx->d_ttiles_z); dfree(ctx->d_bimg_z);
    CU(cudaMalloc(&ctx->d_w, sizeof(float4) * (sum_len + 80)));       // slack: the rescorer never reads past len, but keep loads in bounds
    CU(cudaMemset(ctx->d_w, 0, sizeof(float4) * (sum_len + 80)));
    CU(cudaMalloc(&ctx->d_woff, 4 * n_cols)); CU(cudaMalloc(&ctx->d_len, 4 * n_cols));
    CU(cudaMalloc(&ctx->d_orig, 4 * n_cols)); CU(cudaMalloc(&ctx->d_thr, 4 * n_cols));
    CU(cudaMalloc(&ctx->d_gtiles, sizeof(GatherTile) * gt.size()));
    CU(cudaMalloc(&ctx->d_ttiles, sizeof(TcTile) * tt.size()));
    CU(cudaMalloc(&ctx->d_bimg, bimg.size() + 128));
    CU(cudaMalloc(&ctx->d_ttiles_z, sizeof(TcTile) * tt_z.size()));
    CU(cudaMalloc(&ctx->d_bimg_z, bimg_z.size() + 128));
    CU(cudaMemcpy(ctx->d_w, w.data(), sizeof(float4) * sum_len, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_woff, woff.data(), 4 * n_cols, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_len, len.data(), 4 * n_cols, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_orig, orig.data(), 4 * n_cols, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_thr, thr_s.data(), 4 * n_cols, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_gtiles, gt.data(), sizeof(GatherTile) * gt.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_ttiles, tt.data(), sizeof(TcTile) * tt.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_bimg, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_ttiles_z, tt_z.data(), sizeof(TcTile) * tt_z.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_bimg_z, bimg_z.data(), bimg_z.size(), cudaMemcpyHostToDevice));
    ctx->gtiles = gt; ctx->ttiles = tt;
    {   // fused rescorer: shared memory for the largest tile's FP32 weights (both tilings), capped; room for the partial raw
        // blocks the filter's warps close at every change of column tile
        // (tiles of 256 padded columns with packed accumulators tag their raw blocks per 128-column half, filter_tc.cuh: halfTags)
        uint32_t mw = 1;
        for (const auto* tv : {&tt, &tt_z})
            for (const TcTile& t : *tv) {
                const bool halves = TC_HALF_TAGS && t.acc16 && t.n_pad == 256;
                uint32_t nw = 0, nw_half = 0;
                for (uint32_t c = t.col0; c < t.col0 + t.n_cols; c++) {
                    nw += len[c];
                    if (c - t.col0 == 127) { nw_half = nw; nw = halves ? 0 : nw; }
                }
                mw = std::max(mw, std::max(nw, halves ? nw_half : 0u));
            }
        ctx->fuse_max_w = std::min(mw, kFuseMaxW);
        if (const char* e = getenv("B200SCAN_FUSE_MAXW")) ctx->fuse_max_w = std::min<uint32_t>(ctx->fuse_max_w, (uint32_t)std::max(1, atoi(e)));     // experiment: smaller tables, more CTAs per SM
        int per_sm = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rescore_tile_kernel<false>, (int)kFuseThreads, fuse_smem_bytes(ctx->fuse_max_w)));
        ctx->fuse_ctas_per_sm = std::max(1, std::min(per_sm, 4));
        const unsigned long long want = ctx->cand_cap / kRawBlock + 4096 + (unsigned long long)ctx->sm_count * kTcEpiWarps * (tt.size() + 2);
        if (want > ctx->blk_cap) {
            dfree(ctx->d_raw); dfree(ctx->d_blk_count); dfree(ctx->d_blk_tag);
            ctx->blk_cap = (uint32_t)want;
            CU(cudaMalloc(&ctx->d_raw, ((size_t)ctx->blk_cap + 1) * kRawBlock * kRawWords * 4));
            CU(cudaMalloc(&ctx->d_blk_count, (size_t)ctx->blk_cap * 4));
            CU(cudaMalloc(&ctx->d_blk_tag, (size_t)ctx->blk_cap * 4));
        }
    }
    ctx->h_len_sorted = len; ctx->h_woff_sorted = woff; ctx->h_orig_sorted = orig;
    ctx->hist_bins = 0;
    ctx->n_cols = (uint32_t)n_cols; ctx->max_len = max_len; ctx->sum_len = sum_len;
    ctx->tc_usable = tc_ok;
    ctx->have_motifs = true;
    for (auto& s : ctx->slot) s.resident = false;
    return B200SCAN_OK;
}

MotifDev motif_dev(const b200scan_ctx* c) { return MotifDev{c->d_w, c->d_woff, c->d_len, c->d_thr, c->d_orig, c->n_cols}; }

BlockDev block_dev(const Slot& s)
{
    BlockDev b;
    b.codes = s.d_codes; b.zmask = s.d_zmask; b.frag = s.d_frag;
    b.has_zero = reinterpret_cast<const uint32_t*>(s.d_counters + 2);
    b.n_total = (uint32_t)s.n_total; b.n_payload = (uint32_t)s.n_payload; b.n_frag = (uint32_t)s.n_frag;
    return b;
}

// Launch the scoring kernels for the block resident in `s` (counters must already be reset).
// Returns the number of kernels launched; records ev[3] after the dominant kernel(s) and ev[4] after the rescorer.
int ensure_order(b200scan_ctx* ctx, Slot& s);

int launch_scoring(b200scan_ctx* ctx, Slot& s, cudaEvent_t ev_after_score, cudaEvent_t ev_after_rescore, int* launches, cudaEvent_t ev_after_order = nullptr)
{
    const MotifDev md = motif_dev(ctx);
    const BlockDev blk = block_dev(s);
    const bool ordered = s.hit_bytes == B200SCAN_HITS_8;
    const uint32_t n_fine = (uint32_t)((s.n_payload + kBucketSize - 1) >> kBucketShift);
    const uint32_t n_coarse = (uint32_t)((s.n_payload + kCoarseSize - 1) >> kCoarseShift);
    if (ordered) {                                   // per-coarse-bucket counters of this block (rescore / gather count into them)
        int rc = ensure_order(ctx, s);
        if (rc) return rc;
        if (n_coarse) CU(cudaMemsetAsync(ctx->d_bucket_cnt, 0, (size_t)n_coarse * 4, ctx->stream));
    }
    HitSink sink{s.d_hits, s.d_counters + 1, s.hit_cap, s.hit_bytes == B200SCAN_HITS_16 ? 0u : 1u, ordered ? ctx->d_bucket_cnt : nullptr};
    unsigned int* err = reinterpret_cast<unsigned int*>(s.d_counters + 2) + 1;
    unsigned int* work = reinterpret_cast<unsigned int*>(s.d_counters + 3);
    unsigned int* work2 = reinterpret_cast<unsigned int*>(s.d_counters + 4);      // work counter of the FP32-accumulator instance
    int n = 0;
    if (s.n_payload == 0) {                          // nothing to score (empty block): keep the event protocol intact
        if (ev_after_score) CU(cudaEventRecord(ev_after_score, ctx->stream));
        if (ev_after_rescore) CU(cudaEventRecord(ev_after_rescore, ctx->stream));
        if (ordered) CU(cudaMemsetAsync(s.d_bucket_start, 0, 4, ctx->stream));
        if (ev_after_order) CU(cudaEventRecord(ev_after_order, ctx->stream));
        return B200SCAN_OK;
    }
    const bool want_tc = (ctx->engine != B200SCAN_ENGINE_GATHER) && ctx->tc_usable;
    const dim3 ggrid((unsigned)((s.n_payload + kGatherSpan - 1) / kGatherSpan), (unsigned)ctx->gtiles.size());
    if (want_tc) {
        TcParams tp;
        tp.work_counter = work; tp.raw = ctx->d_raw; tp.blk_count = ctx->d_blk_count; tp.blk_tag = ctx->d_blk_tag; tp.n_blocks = work + 1; tp.blk_cap = ctx->blk_cap;
        tp.error_flag = err;
        tp.trace = ctx->d_trace;
        // One instance per accumulator type over its share of the tiles (FP16 tiles are stored first), and that for both
        // kinds of block: plain ACGT, or with zero-contribution characters (the instance that does not match the block's
        // has_zero flag returns at once).  All append to the same raw-entry blocks; each accumulator type has its own work counter.
        const uint32_t n_tiles = (uint32_t)ctx->ttiles.size();
        using FilterFn = void (*)(TcParams, BlockDev);
        static const FilterFn kFilter[2][2][2] = {       // [FP16 accumulators][masked block][CTA pair]
            {{filter_tc_kernel<false, false, false>, filter_tc_kernel<false, false, true>}, {filter_tc_kernel<false, true, false>, filter_tc_kernel<false, true, true>}},
            {{filter_tc_kernel<true, false, false>, filter_tc_kernel<true, false, true>}, {filter_tc_kernel<true, true, false>, filter_tc_kernel<true, true, true>}}};
        const int pair = ctx->pair_mode ? 1 : 0;
        tp.n_spans = (uint32_t)((s.n_payload + (pair ? 2 : 1) * (uint64_t)kTcSpan - 1) / ((pair ? 2 : 1) * (uint64_t)kTcSpan));
        static const FilterFn kFilter8[2] = {filter_tc_kernel<true, false, false, true>, filter_tc_kernel<true, true, false, true>};     // INT8 operands [masked block]
        unsigned int* work3 = work2 + 1;                                               // work counter of the INT8 instance
        auto launch = [&](int acc16, int z) -> cudaError_t {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(pair ? 2u * (unsigned)(ctx->sm_count / 2) : (unsigned)(ctx->sm_count * kTcCtasPerSm));
            cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kTcSmemBytes; cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = pair ? 1 : 0;
            if (acc16 == 2) {                                  // INT8 instances are single-CTA
                cfg.gridDim = dim3((unsigned)(ctx->sm_count * kTcCtasPerSm)); cfg.numAttrs = 0;
                return cudaLaunchKernelEx(&cfg, kFilter8[z], tp, blk);
            }
            return cudaLaunchKernelEx(&cfg, kFilter[acc16][z][pair], tp, blk);
        };
        for (int z = 0; z < 2; z++) {
            const TcTile* tiles = z ? ctx->d_ttiles_z : ctx->d_ttiles;
            const uint32_t n8 = z ? ctx->n_tiles8_z : ctx->n_tiles8, n16 = z ? ctx->n_tiles16_z : ctx->n_tiles16;
            tp.bimg = z ? ctx->d_bimg_z : ctx->d_bimg;
            if (n8) {
                // (a pair launch counts spans of 2 * kTcSpan windows; the INT8 instance is single-CTA)
                const uint32_t spans = tp.n_spans;
                tp.n_spans = (uint32_t)((s.n_payload + (uint64_t)kTcSpan - 1) / (uint64_t)kTcSpan);
                tp.tiles = tiles; tp.n_tiles = n8; tp.work_counter = work3;
                CU(launch(2, z));
                tp.n_spans = spans;
                n++;
            }
            if (n16) {
                tp.tiles = tiles + n8; tp.n_tiles = n16; tp.work_counter = work;
                CU(launch(1, z));
                n++;
            }
            if (n8 + n16 < n_tiles) {
                tp.tiles = tiles + n8 + n16; tp.n_tiles = n_tiles - n8 - n16; tp.work_counter = work2;
                CU(launch(0, z));
                n++;
            }
        }
        if (ev_after_score) CU(cudaEventRecord(ev_after_score, ctx->stream));
        {
            // expand + exact rescoring in one kernel, a column tile's weights at a time in shared memory (the instance that does not
            // match the block's has_zero flag returns at once)
            unsigned int* fwork = reinterpret_cast<unsigned int*>(s.d_counters + 5);
            const size_t fsm = fuse_smem_bytes(ctx->fuse_max_w);
            rescore_tile_kernel<false><<<ctx->sm_count * ctx->fuse_ctas_per_sm, kFuseThreads, fsm, ctx->stream>>>(md, blk, ctx->d_raw, ctx->d_blk_count, ctx->d_blk_tag, work + 1, ctx->blk_cap,
                                                                                        fwork, s.d_counters, ctx->fuse_max_w, sink);
            rescore_tile_kernel<true><<<ctx->sm_count * ctx->fuse_ctas_per_sm, kFuseThreads, fsm, ctx->stream>>>(md, blk, ctx->d_raw, ctx->d_blk_count, ctx->d_blk_tag, work + 1, ctx->blk_cap,
                                                                                       fwork, s.d_counters, ctx->fuse_max_w, sink);
            n += 2;
        }
        if (ev_after_rescore) CU(cudaEventRecord(ev_after_rescore, ctx->stream));
    } else {
        gather_scan_kernel<false><<<ggrid, kGatherThreads, ctx->gather_smem, ctx->stream>>>(md, blk, ctx->d_gtiles, sink, 0);
        gather_scan_kernel<true><<<ggrid, kGatherThreads, ctx->gather_smem, ctx->stream>>>(md, blk, ctx->d_gtiles, sink, 1);
        n += 2;
        if (ev_after_score) CU(cudaEventRecord(ev_after_score, ctx->stream));
        if (ev_after_rescore) CU(cudaEventRecord(ev_after_rescore, ctx->stream));
    }
    if (ordered) {
        // (position, column) order on the device: scan of the per-bucket counts, scatter into bucket-contiguous order, order
        // inside every bucket; the final 8-byte records overwrite the slot's (now dead) unordered list.
        bucket_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_bucket_cnt, n_coarse, ctx->d_coarse_start, ctx->d_bucket_cursor);
        bucket_scatter_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reinterpret_cast<const Hit12*>(s.d_hits), s.d_counters + 1, s.hit_cap,
                                                                         ctx->d_bucket_cursor, ctx->d_sort_tmp);
        bucket_order_kernel<<<ctx->sm_count * 4, kOrderThreads, 0, ctx->stream>>>(ctx->d_sort_tmp, ctx->d_coarse_start, n_coarse, n_fine, s.d_bucket_start,
                                                                                 reinterpret_cast<uint2*>(s.d_hits));
        n += 3;
    }
    if (ev_after_order) CU(cudaEventRecord(ev_after_order, ctx->stream));
    CU(cudaGetLastError());
    if (launches) *launches += n;
    return B200SCAN_OK;
}

// ---- lazily allocated buffers ----------------------------------------------------------------------------
// device block buffers + hit list of a slot (first submit on the slot)
int ensure_slot(b200scan_ctx* ctx, Slot& s)
{
    if (s.d_codes) return B200SCAN_OK;
    const size_t nb = (size_t)ctx->max_block;
    CU(cudaMalloc(&s.d_codes, nb / 4 + kPadBytes));
    CU(cudaMalloc(&s.d_zmask, nb / 8 + kPadBytes));
    CU(cudaMemsetAsync(s.d_codes, 0, nb / 4 + kPadBytes, ctx->up_stream));
    CU(cudaMemsetAsync(s.d_zmask, 0, nb / 8 + kPadBytes, ctx->up_stream));
    CU(cudaMalloc(&s.d_hits, sizeof(b200scan_hit) * ctx->max_hits));
    s.hit_cap = ctx->max_hits;
    return B200SCAN_OK;
}
// character staging (b200scan_submit_ascii / b200scan_hist_block_ascii only)
int ensure_ascii(b200scan_ctx* ctx, Slot& s, bool need_host)
{
    const size_t nb = (size_t)ctx->max_block;
    if (!s.d_ascii) CU(cudaMalloc(&s.d_ascii, nb + kPadBytes));
    if (need_host && !s.h_ascii) CU(cudaMallocHost(&s.h_ascii, nb + 64));
    return B200SCAN_OK;
}
// B200SCAN_HITS_8: bucket arrays for max_block characters, scratch list for the slot's hit capacity
int ensure_order(b200scan_ctx* ctx, Slot& s)
{
    const size_t want = (size_t)(ctx->max_block >> kBucketShift) + 2, want_c = (size_t)(ctx->max_block >> kCoarseShift) + 2;
    if (ctx->bucket_cap < want_c) {
        dfree(ctx->d_bucket_cnt); dfree(ctx->d_bucket_cursor); dfree(ctx->d_coarse_start);
        CU(cudaMalloc(&ctx->d_bucket_cnt, want_c * 4)); CU(cudaMalloc(&ctx->d_bucket_cursor, want_c * 4)); CU(cudaMalloc(&ctx->d_coarse_start, want_c * 4));
        ctx->bucket_cap = want_c;
    }
    if (s.bucket_cap < want) {
        dfree(s.d_bucket_start);
        CU(cudaMalloc(&s.d_bucket_start, want * 4));
        s.bucket_cap = want;
    }
    if (ctx->sort_cap < s.hit_cap) {
        CU(cudaStreamSynchronize(ctx->stream));          // a previous block's ordering kernels may still read the old scratch list
        dfree(ctx->d_sort_tmp);
        ctx->sort_cap = 0;
        CU(cudaMalloc(&ctx->d_sort_tmp, sizeof(Hit12) * s.hit_cap));
        ctx->sort_cap = s.hit_cap;
    }
    return B200SCAN_OK;
}
// pinned host copy of a collected hit list: sized from the lists themselves
int ensure_host_hits(b200scan_ctx* ctx, Slot& s, size_t bytes)
{
    if (bytes <= s.h_hit_bytes) return B200SCAN_OK;
    hfree(s.h_hits);
    s.h_hit_bytes = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    CU(cudaMallocHost(&s.h_hits, want));
    s.h_hit_bytes = want;
    return B200SCAN_OK;
}

int reset_counters(b200scan_ctx* ctx, Slot& s, bool keep_has_zero, cudaStream_t st)
{
    // [0] n_cand [1] n_hits: zero.  [2] = {has_zero, error}: keep has_zero on re-runs.  [3] {work counter, raw blocks} [4] second work counter: zero.
    // (all on the caller's stream `st`: on a first submit that is the upload stream, where the pack kernel then sets has_zero)
    CU(cudaMemsetAsync(s.d_counters, 0, 16, st));
    if (keep_has_zero) CU(cudaMemsetAsync(reinterpret_cast<uint32_t*>(s.d_counters + 2) + 1, 0, 4, st));
    else CU(cudaMemsetAsync(s.d_counters + 2, 0, 8, st));
    CU(cudaMemsetAsync(s.d_counters + 3, 0, 24, st));               // the work counters (filter instances, fused rescorer) and the block counter
    return B200SCAN_OK;
}

int check_common(b200scan_ctx* ctx, int slot, uint64_t n_total, uint64_t n_payload, const uint64_t* frag, uint64_t n_frag)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (slot < 0 || slot >= B200SCAN_NUM_SLOTS) return fail(ctx, B200SCAN_EINVAL, "slot %d out of range", slot);
    if (!ctx->have_motifs) return fail(ctx, B200SCAN_ESTATE, "b200scan_set_motifs has not been called");
    if (ctx->slot[slot].in_flight) return fail(ctx, B200SCAN_ESTATE, "slot %d has an uncollected block", slot);
    if (n_payload > n_total) return fail(ctx, B200SCAN_EINVAL, "n_payload > n_total");
    if (n_total > ctx->max_block) return fail(ctx, B200SCAN_ELIMIT, "block of %llu characters exceeds max_block_nt %llu",
                                             (unsigned long long)n_total, (unsigned long long)ctx->max_block);
    if (n_frag && !frag) return fail(ctx, B200SCAN_EINVAL, "frag_starts is NULL");
    for (uint64_t i = 0; i < n_frag; i++)
        if (frag[i] == 0 || frag[i] >= n_total || (i && frag[i] <= frag[i - 1]))
            return fail(ctx, B200SCAN_EINVAL, "frag_starts must be strictly ascending inside (0, n_total)");
    return B200SCAN_OK;
}

int stage_frags(b200scan_ctx* ctx, Slot& s, const uint64_t* frag, uint64_t n_frag, cudaStream_t st)
{
    if (n_frag > s.frag_cap) {
        size_t cap = std::max<size_t>(n_frag * 2, 1 << 16);
        hfree(s.h_frag); dfree(s.d_frag);
        CU(cudaMallocHost(&s.h_frag, cap * 4)); CU(cudaMalloc(&s.d_frag, cap * 4));
        s.frag_cap = cap;
    }
    for (uint64_t i = 0; i < n_frag; i++) s.h_frag[i] = (uint32_t)frag[i];
    if (n_frag) CU(cudaMemcpyAsync(s.d_frag, s.h_frag, n_frag * 4, cudaMemcpyHostToDevice, st));
    return B200SCAN_OK;
}

int finish_submit(b200scan_ctx* ctx, Slot& s)
{
    s.timing.kernel_launches = 0;
    int launches = 0;
    // the block was uploaded and packed on the upload stream (behind the kernels of the other slot's block)
    CU(cudaStreamWaitEvent(ctx->stream, s.ev[2], 0));
    CU(cudaEventRecord(s.ev[8], ctx->stream));
    int rc = launch_scoring(ctx, s, s.ev[3], s.ev[4], &launches, s.ev[9]);
    if (rc) return rc;
    s.timing.kernel_launches += launches;
    CU(cudaMemcpyAsync(s.h_counters, s.d_counters, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(s.ev[5], ctx->stream));
    s.in_flight = true; s.resident = true;
    return B200SCAN_OK;
}

} // namespace

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

int b200scan_abi_version(void) { return B200SCAN_ABI_VERSION; }

int b200scan_device_count(void)
{
    int n = 0, ok = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

void* b200scan_host_alloc(uint64_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void b200scan_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* b200scan_last_error(const b200scan_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b200scan_create(b200scan_ctx** out, int device, uint64_t max_block_nt, uint64_t max_hits)
{
    b200scan_ctx* ctx = nullptr;      // CU() reports into g_create_error while ctx == nullptr
    if (!out) return B200SCAN_EINVAL;
    *out = nullptr;
    if (max_block_nt == 0 || max_block_nt > 0xF0000000ull) return fail(nullptr, B200SCAN_EINVAL, "max_block_nt must be in (0, 2^32)");
    // B200SCAN_TIMING=1: wall-clock phases of the creation on stderr (CUDA start-up dominates short runs)
    const bool timing = getenv("B200SCAN_TIMING") != nullptr;
    auto tPrev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200scan_create] %-34s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t - tPrev).count());
        tPrev = t;
    };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, B200SCAN_ENODEVICE, "no CUDA device (this library has no CPU path)");
    lap("cudaGetDeviceCount");
    if (device < 0 || device >= ndev) return fail(nullptr, B200SCAN_ENODEVICE, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(nullptr, B200SCAN_ENODEVICE, "device %d is sm_%d%d; this build is sm_100a only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    CU(cudaFree(nullptr));
    lap("device properties + primary context");

    b200scan_ctx* c = new b200scan_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount; c->max_block = max_block_nt;
    ctx = c;
    auto bail = [&](int rc) { std::string e = c->err; b200scan_destroy(c); g_create_error = e; return rc; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, B200SCAN_ECUDA, "%s -> %s", #call, cudaGetErrorString(e_)); \
                       return bail(e_ == cudaErrorMemoryAllocation ? B200SCAN_ENOMEM : B200SCAN_ECUDA); } } while (0)
    CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    if (const char* e = getenv("B200SCAN_PAIR")) c->pair_mode = atoi(e) != 0;
    if (const char* e = getenv("B200SCAN_MARGIN16_SCALE")) c->margin16_scale = atof(e);
    if (const char* e = getenv("B200SCAN_I8_MAX_OVERSHOOT")) c->i8_max_overshoot = atof(e);
    if (const char* e = getenv("B200SCAN_HIST_KERNEL")) c->hist_kernel = std::max(0, std::min(3, atoi(e)));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(filter_tc_kernel<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    CUB(cudaFuncSetAttribute(gather_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kGatherSmemW + 16384)));
    CUB(cudaFuncSetAttribute(gather_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kGatherSmemW + 16384)));
    lap("streams + kernel attributes");
    if (max_hits == 0) max_hits = 1 << 20;
    {
        // Budget for the buffers that grow with the hit density: about 96 bytes of device memory per hit record of the densest
        // block (three slots x 16 B hit list, 12 B ordering scratch, 2 x 16 B raw entries for the ~1.2 candidates per hit)
        // may take 60 % of what is free now.  A block that needs more is refused with B200SCAN_ENOMEM (the caller submits it in
        // smaller pieces: the CLI halves it, cli.cpp: scanSplit); B200SCAN_HIT_BUDGET=<records> overrides the figure.
        size_t free_b = 0, total_b = 0;
        CUB(cudaMemGetInfo(&free_b, &total_b));
        c->hit_budget = std::max<unsigned long long>((unsigned long long)(0.6 * (double)free_b / 96.0), 1 << 16);
        if (const char* e = getenv("B200SCAN_HIT_BUDGET")) c->hit_budget = std::max<unsigned long long>(strtoull(e, nullptr, 10), 1024);
        max_hits = std::min<unsigned long long>(max_hits, c->hit_budget);
    }
    c->max_hits = max_hits;
    // Per slot only the counters and events exist from the start; block buffers, hit lists and pinned staging memory are
    // allocated at the slot's first use (ensure_slot / ensure_ascii / ensure_host_hits): a short run does not pay ~0.5 s for
    // pinning memory it never touches.
    for (auto& s : c->slot) {
        CUB(cudaMallocHost(&s.h_counters, 64));
        CUB(cudaMalloc(&s.d_counters, 64));
        CUB(cudaMemset(s.d_counters, 0, 64));
        for (auto& e : s.ev) CUB(cudaEventCreate(&e));
    }
    lap("counters + events of the slots");
#if defined(B200_TRACE) || defined(B200_PHASE)
    CUB(cudaMalloc(&c->d_trace, 4 * kTraceTiles * 4 * 8));
    CUB(cudaMemset(c->d_trace, 0, 4 * kTraceTiles * 4 * 8));
#endif
    c->cand_cap = std::max<unsigned long long>(2 * max_hits, 1 << 20);
    c->blk_cap = (uint32_t)std::max<unsigned long long>(c->cand_cap / kRawBlock + 4096, 8192);
    CUB(cudaFuncSetAttribute(rescore_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fuse_smem_bytes(kFuseMaxW)));
    CUB(cudaFuncSetAttribute(rescore_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fuse_smem_bytes(kFuseMaxW)));
    CUB(cudaMalloc(&c->d_raw, ((size_t)c->blk_cap + 1) * kRawBlock * kRawWords * 4));     // + the sacrificial overflow block
    CUB(cudaMalloc(&c->d_blk_count, (size_t)c->blk_cap * 4));
    CUB(cudaMalloc(&c->d_blk_tag, (size_t)c->blk_cap * 4));
    lap("candidate / raw buffers");
#undef CUB
    *out = c;
    return B200SCAN_OK;
}

void b200scan_destroy(b200scan_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& s : c->slot) {
        hfree(s.h_ascii); hfree(s.h_frag); hfree(s.h_hits); hfree(s.h_counters); hfree(s.h_bucket);
        dfree(s.d_ascii); dfree(s.d_codes); dfree(s.d_zmask); dfree(s.d_frag); dfree(s.d_hits); dfree(s.d_counters); dfree(s.d_bucket_start);
        for (auto& e : s.ev) if (e) cudaEventDestroy(e);
    }
    dfree(c->d_raw); dfree(c->d_blk_count); dfree(c->d_blk_tag); dfree(c->d_flush);
    dfree(c->d_bucket_cnt); dfree(c->d_bucket_cursor); dfree(c->d_coarse_start); dfree(c->d_sort_tmp);
    dfree(c->d_w); dfree(c->d_woff); dfree(c->d_len); dfree(c->d_orig); dfree(c->d_thr);
    dfree(c->d_gtiles); dfree(c->d_ttiles); dfree(c->d_bimg); dfree(c->d_ttiles_z); dfree(c->d_bimg_z);
    dfree(c->d_hist); dfree(c->d_hmin); dfree(c->d_hwid); dfree(c->d_htiles); dfree(c->d_htiles2);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->up_stream) { cudaStreamSynchronize(c->up_stream); cudaStreamDestroy(c->up_stream); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int b200scan_set_engine(b200scan_ctx* ctx, int engine)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (engine < B200SCAN_ENGINE_AUTO || engine > B200SCAN_ENGINE_TENSOR) return fail(ctx, B200SCAN_EINVAL, "unknown engine %d", engine);
    ctx->engine = engine;
    return B200SCAN_OK;
}

int b200scan_set_tensor_accumulator(b200scan_ctx* ctx, int bits)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (bits != 0 && bits != 8 && bits != 16 && bits != 32) return fail(ctx, B200SCAN_EINVAL, "accumulator must be 0 (auto), 8 (INT8 operands), 16 or 32");
    ctx->acc_pref = bits;
    return B200SCAN_OK;
}

int b200scan_set_motifs(b200scan_ctx* ctx, const float* P, int32_t ldp, int32_t n_cols, const int32_t* col_len, const float* thr)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (!P || !col_len || !thr || n_cols <= 0) return fail(ctx, B200SCAN_EINVAL, "set_motifs: NULL or empty input");
    for (auto& s : ctx->slot) if (s.in_flight) return fail(ctx, B200SCAN_ESTATE, "set_motifs while a block is in flight");
    for (int32_t c = 0; c < n_cols; c++) {
        if (col_len[c] < 1) return fail(ctx, B200SCAN_EINVAL, "column %d has length %d", c, col_len[c]);
        if (col_len[c] > B200SCAN_MAX_MOTIF_LEN) return fail(ctx, B200SCAN_ELIMIT, "column %d has length %d > %d", c, col_len[c], B200SCAN_MAX_MOTIF_LEN);
        if (4 * col_len[c] > ldp) return fail(ctx, B200SCAN_EINVAL, "ldp %d < 4*len of column %d", ldp, c);
    }
    if (ctx->hit_format == B200SCAN_HITS_8 && (uint32_t)n_cols > (1u << 24))
        return fail(ctx, B200SCAN_ELIMIT, "B200SCAN_HITS_8 records hold 24-bit column indices (%d columns)", n_cols);
    NvtxRange nvtx("b200scan_set_motifs");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return build_motifs(ctx, P, ldp, n_cols, col_len, thr);
}

int b200scan_submit_ascii(b200scan_ctx* ctx, int slot, const char* block, uint64_t n_total, uint64_t n_payload,
                          const uint64_t* frag_starts, uint64_t n_frag, int lowercase_mode)
{
    NvtxRange nvtx("b200scan_submit_ascii");
    int rc = check_common(ctx, slot, n_total, n_payload, frag_starts, n_frag);
    if (rc) return rc;
    if (!block && n_total) return fail(ctx, B200SCAN_EINVAL, "block is NULL");
    if (lowercase_mode != B200SCAN_LOWER_ZERO && lowercase_mode != B200SCAN_LOWER_FOLD) return fail(ctx, B200SCAN_EINVAL, "bad lowercase_mode");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[slot];
    s.n_total = n_total; s.n_payload = n_payload; s.n_frag = n_frag;
    s.timing = b200scan_timing{};
    s.hit_bytes = ctx->hit_format;
    // pinned caller memory goes straight to the device; pageable memory is staged through the slot's pinned buffer
    const void* src = block;
    cudaPointerAttributes pa;
    const bool pageable = n_total && (cudaPointerGetAttributes(&pa, block) != cudaSuccess || pa.type != cudaMemoryTypeHost);
    cudaGetLastError();
    rc = ensure_slot(ctx, s);
    if (rc == B200SCAN_OK) rc = ensure_ascii(ctx, s, pageable);
    if (rc) return rc;
    CU(cudaEventRecord(s.ev[0], ctx->up_stream));
    if (pageable) {
        std::memcpy(s.h_ascii, block, n_total);
        src = s.h_ascii;
    }
    if (n_total) CU(cudaMemcpyAsync(s.d_ascii, src, n_total, cudaMemcpyHostToDevice, ctx->up_stream));
    rc = stage_frags(ctx, s, frag_starts, n_frag, ctx->up_stream);
    if (rc) return rc;
    rc = reset_counters(ctx, s, false, ctx->up_stream);
    if (rc) return rc;
    CU(cudaEventRecord(s.ev[1], ctx->up_stream));
    if (n_total) {
        const unsigned threads = (unsigned)((n_total + 31) / 32);
        pack_ascii_kernel<<<(threads + 255) / 256, 256, 0, ctx->up_stream>>>(s.d_ascii, (uint32_t)n_total, lowercase_mode == B200SCAN_LOWER_FOLD,
                                                                          s.d_codes, s.d_zmask, reinterpret_cast<uint32_t*>(s.d_counters + 2));
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(s.ev[2], ctx->up_stream));
    rc = finish_submit(ctx, s);
    if (rc == B200SCAN_OK && n_total) s.timing.kernel_launches += 1;
    return rc;
}

int b200scan_submit_packed(b200scan_ctx* ctx, int slot, const uint32_t* codes2, const uint32_t* zero_mask, uint64_t n_total,
                           uint64_t n_payload, const uint64_t* frag_starts, uint64_t n_frag)
{
    NvtxRange nvtx("b200scan_submit_packed");
    int rc = check_common(ctx, slot, n_total, n_payload, frag_starts, n_frag);
    if (rc) return rc;
    if (!codes2 && n_total) return fail(ctx, B200SCAN_EINVAL, "codes2 is NULL");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[slot];
    s.n_total = n_total; s.n_payload = n_payload; s.n_frag = n_frag;
    s.timing = b200scan_timing{};
    s.hit_bytes = ctx->hit_format;
    rc = ensure_slot(ctx, s);
    if (rc) return rc;
    CU(cudaEventRecord(s.ev[0], ctx->up_stream));
    const size_t cw = (n_total + 15) / 16, zw = (n_total + 31) / 32;
    // pageable sources: cudaMemcpyAsync stages them itself and returns once the source may be reused
    if (cw) CU(cudaMemcpyAsync(s.d_codes, codes2, cw * 4, cudaMemcpyHostToDevice, ctx->up_stream));
    uint32_t hz = 0;
    if (zero_mask) {
        for (size_t i = 0; i < zw && !hz; i++) {
            uint32_t live = (i + 1 == zw && (n_total & 31)) ? ((1u << (n_total & 31)) - 1u) : 0xffffffffu;
            if (zero_mask[i] & live) hz = 1;
        }
        if (hz) CU(cudaMemcpyAsync(s.d_zmask, zero_mask, zw * 4, cudaMemcpyHostToDevice, ctx->up_stream));
    }
    rc = stage_frags(ctx, s, frag_starts, n_frag, ctx->up_stream);
    if (rc) return rc;
    rc = reset_counters(ctx, s, false, ctx->up_stream);
    if (rc) return rc;
    if (hz) CU(cudaMemsetAsync(s.d_counters + 2, 1, 1, ctx->up_stream));      // has_zero = 1 (little endian)
    CU(cudaEventRecord(s.ev[1], ctx->up_stream));
    CU(cudaEventRecord(s.ev[2], ctx->up_stream));
    return finish_submit(ctx, s);
}

int b200scan_hist_begin(b200scan_ctx* ctx, const float* col_min, const float* col_max, uint32_t num_bins)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (!ctx->have_motifs) return fail(ctx, B200SCAN_ESTATE, "b200scan_set_motifs has not been called");
    if (!col_min || !col_max || num_bins == 0) return fail(ctx, B200SCAN_EINVAL, "hist_begin: NULL or empty input");
    for (auto& sl : ctx->slot) if (sl.in_flight) return fail(ctx, B200SCAN_ESTATE, "hist_begin while a block is in flight");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const uint32_t n = ctx->n_cols;
    // column tiles: FP32 weights + metadata + u32 histograms of the tile must fit in shared memory
    const size_t budget = 160 * 1024;
    std::vector<GatherTile> ht;
    size_t max_smem = 0;
    for (uint32_t sc = 0; sc < n;) {
        GatherTile t{sc, 0, ctx->h_woff_sorted[sc], 0};
        auto need = [&](uint32_t nw, uint32_t nc) { return (size_t)nw * 16 + (size_t)nc * (16 + (size_t)num_bins * 4); };
        while (sc < n && t.n_cols < 256 && need(t.n_w + ctx->h_len_sorted[sc], t.n_cols + 1) <= budget) {
            t.n_w += ctx->h_len_sorted[sc]; t.n_cols++; sc++;
        }
        if (t.n_cols == 0) return fail(ctx, B200SCAN_ELIMIT, "num_bins %u too large for the shared-memory histograms", num_bins);
        max_smem = std::max(max_smem, need(t.n_w, t.n_cols));
        ht.push_back(t);
    }
    std::vector<float> mn(n), wid(n);
    for (uint32_t sc = 0; sc < n; sc++) {
        const uint32_t c = ctx->h_orig_sorted[sc];
        mn[sc] = col_min[c];
        wid[sc] = (col_max[c] - col_min[c]) / (float)num_bins;        // ScoreHistogram ctor, motif.h:66
    }
    // the same for gather_hist2_kernel: position rows of the tile's longest column (+ the zero row) for every column, padded to
    // groups of kHist2U columns; 100 KB per CTA so that two CTAs share an SM (a column whose table alone exceeds that gets a
    // tile of its own within the large budget)
    std::vector<HistTile2> ht2;
    if (ctx->hist_kernel >= 2) {
        const size_t budget2 = 100 * 1024;
        for (uint32_t sc = 0; sc < n;) {
            HistTile2 t{sc, 0, 0, 0};
            while (sc < n && t.n_cols < 248 && hist2_smem_bytes(t.n_cols + 1, ctx->h_len_sorted[sc], num_bins) <= (t.n_cols ? budget2 : budget)) {
                t.n_cols++; t.max_len = ctx->h_len_sorted[sc]; sc++;
            }
            if (t.n_cols == 0) { ht2.clear(); break; }      // a table this large only fits the narrower tiles of gather_hist_kernel: that kernel runs instead
            t.n_pad = (t.n_cols + kHist2U - 1) / kHist2U * kHist2U;
            ht2.push_back(t);
        }
        for (const HistTile2& t : ht2) max_smem = std::max(max_smem, hist2_smem_bytes(t.n_cols, t.max_len, num_bins));
    }
    dfree(ctx->d_hist); dfree(ctx->d_hmin); dfree(ctx->d_hwid); dfree(ctx->d_htiles); dfree(ctx->d_htiles2);
    CU(cudaMalloc(&ctx->d_hist, (size_t)n * num_bins * 8));
    CU(cudaMemset(ctx->d_hist, 0, (size_t)n * num_bins * 8));
    CU(cudaMalloc(&ctx->d_hmin, 4 * n)); CU(cudaMalloc(&ctx->d_hwid, 4 * n));
    CU(cudaMalloc(&ctx->d_htiles, sizeof(GatherTile) * ht.size()));
    CU(cudaMemcpy(ctx->d_hmin, mn.data(), 4 * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_hwid, wid.data(), 4 * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_htiles, ht.data(), sizeof(GatherTile) * ht.size(), cudaMemcpyHostToDevice));
    if (!ht2.empty()) {
        CU(cudaMalloc(&ctx->d_htiles2, sizeof(HistTile2) * ht2.size()));
        CU(cudaMemcpy(ctx->d_htiles2, ht2.data(), sizeof(HistTile2) * ht2.size(), cudaMemcpyHostToDevice));
    }
    CU(cudaFuncSetAttribute(gather_hist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    CU(cudaFuncSetAttribute(gather_hist2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    ctx->htiles = ht; ctx->htiles2 = ht2; ctx->hist_smem = max_smem; ctx->hist_bins = num_bins;
    return B200SCAN_OK;
}

int b200scan_hist_block_ascii(b200scan_ctx* ctx, const char* block, uint64_t n_total, uint64_t n_payload,
                              const uint64_t* frag_starts, uint64_t n_frag, int lowercase_mode)
{
    int rc = check_common(ctx, 0, n_total, n_payload, frag_starts, n_frag);
    if (rc) return rc;
    if (!ctx->hist_bins) return fail(ctx, B200SCAN_ESTATE, "b200scan_hist_begin has not been called");
    if (!block && n_total) return fail(ctx, B200SCAN_EINVAL, "block is NULL");
    if (lowercase_mode != B200SCAN_LOWER_ZERO && lowercase_mode != B200SCAN_LOWER_FOLD) return fail(ctx, B200SCAN_EINVAL, "bad lowercase_mode");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[0];
    CU(cudaStreamSynchronize(ctx->stream));                 // the previous block still reads the staging buffers
    s.n_total = n_total; s.n_payload = n_payload; s.n_frag = n_frag; s.resident = false;
    if (n_total == 0 || n_payload == 0) return B200SCAN_OK;
    rc = ensure_slot(ctx, s);
    if (rc == B200SCAN_OK) rc = ensure_ascii(ctx, s, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->up_stream));              // (ensure_slot clears fresh buffers on the upload stream)
    std::memcpy(s.h_ascii, block, n_total);
    CU(cudaMemcpyAsync(s.d_ascii, s.h_ascii, n_total, cudaMemcpyHostToDevice, ctx->stream));
    rc = stage_frags(ctx, s, frag_starts, n_frag, ctx->stream);
    if (rc) return rc;
    rc = reset_counters(ctx, s, false, ctx->stream);
    if (rc) return rc;
    const unsigned threads = (unsigned)((n_total + 31) / 32);
    pack_ascii_kernel<<<(threads + 255) / 256, 256, 0, ctx->stream>>>(s.d_ascii, (uint32_t)n_total, lowercase_mode == B200SCAN_LOWER_FOLD,
                                                                      s.d_codes, s.d_zmask, reinterpret_cast<uint32_t*>(s.d_counters + 2));
    const MotifDev md = motif_dev(ctx);
    const BlockDev blk = block_dev(s);
    // two instances per block: the one for blocks without zero-contribution characters and the masked one; each returns at once
    // unless the block's has_zero flag (written by the pack kernel just before) asks for it
    if (ctx->hist_kernel >= 2 && !ctx->htiles2.empty()) {
        const dim3 grid((unsigned)((n_payload + kHistSpan - 1) / kHistSpan), (unsigned)ctx->htiles2.size());
        if (ctx->hist_kernel == 3) {
            gather_hist2_kernel<false, true><<<grid, kHist2Threads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles2, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 0);
            gather_hist2_kernel<true, true><<<grid, kHist2Threads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles2, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 1);
        } else {
            gather_hist2_kernel<false, false><<<grid, kHist2Threads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles2, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 0);
            gather_hist2_kernel<true, false><<<grid, kHist2Threads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles2, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 1);
        }
    } else {
        const dim3 grid((unsigned)((n_payload + kHistSpan - 1) / kHistSpan), (unsigned)ctx->htiles.size());
        if (ctx->hist_kernel == 1) {
            gather_hist_kernel<false, true><<<grid, kGatherThreads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 0);
            gather_hist_kernel<true, true><<<grid, kGatherThreads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 1);
        } else {
            gather_hist_kernel<false, false><<<grid, kGatherThreads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 0);
            gather_hist_kernel<true, false><<<grid, kGatherThreads, ctx->hist_smem, ctx->stream>>>(md, blk, ctx->d_htiles, ctx->d_hmin, ctx->d_hwid, ctx->hist_bins, ctx->d_hist, 1);
        }
    }
    CU(cudaGetLastError());
    return B200SCAN_OK;
}

int b200scan_hist_read(b200scan_ctx* ctx, uint64_t* counts, uint64_t n_counts)
{
    if (!ctx || !counts) return B200SCAN_EINVAL;
    if (!ctx->hist_bins) return fail(ctx, B200SCAN_ESTATE, "b200scan_hist_begin has not been called");
    if (n_counts != (uint64_t)ctx->n_cols * ctx->hist_bins) return fail(ctx, B200SCAN_EINVAL, "hist_read: expected %llu counts", (unsigned long long)ctx->n_cols * ctx->hist_bins);
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, B200SCAN_ECUDA, "histogram kernels failed: %s", cudaGetErrorString(e));
    CU(cudaMemcpy(counts, ctx->d_hist, n_counts * 8, cudaMemcpyDeviceToHost));
    return B200SCAN_OK;
}

static int collect_impl(b200scan_ctx* ctx, int slot, int want_bytes, const void** hits, uint64_t* n_hits, b200scan_timing* timing,
                        const uint32_t** bucket_start = nullptr, uint64_t* n_buckets = nullptr)
{
    NvtxRange nvtx("b200scan_collect");
    if (!ctx) return B200SCAN_EINVAL;
    if (slot < 0 || slot >= B200SCAN_NUM_SLOTS) return fail(ctx, B200SCAN_EINVAL, "slot %d out of range", slot);
    Slot& s = ctx->slot[slot];
    if (!s.in_flight) return fail(ctx, B200SCAN_ESTATE, "slot %d has nothing to collect", slot);
    if (s.hit_bytes != want_bytes)
        return fail(ctx, B200SCAN_ESTATE, "slot %d was submitted with %d-byte hit records: collect it with %s", slot, s.hit_bytes,
                    s.hit_bytes == B200SCAN_HITS_12 ? "b200scan_collect12" : (s.hit_bytes == B200SCAN_HITS_8 ? "b200scan_collect8" : "b200scan_collect"));
    CU(cudaSetDevice(ctx->device));
    s.in_flight = false;
    for (int attempt = 0;; attempt++) {
        cudaError_t e = cudaEventSynchronize(s.ev[5]);
        if (e != cudaSuccess) { s.resident = false; return fail(ctx, B200SCAN_ECUDA, "scan failed: %s", cudaGetErrorString(e)); }
        const unsigned long long n_cand = s.h_counters[0], nh = s.h_counters[1];
        const uint32_t has_zero = (uint32_t)(s.h_counters[2] & 0xffffffffu), errflag = (uint32_t)(s.h_counters[2] >> 32);
        if (errflag) return fail(ctx, B200SCAN_ECUDA, "kernel reported error flag 0x%08x", errflag);
        if (ctx->engine == B200SCAN_ENGINE_TENSOR && !ctx->tc_usable)
            return fail(ctx, B200SCAN_ESTATE, "ENGINE_TENSOR unavailable for this motif set");
        const uint32_t n_blocks = (uint32_t)(s.h_counters[3] >> 32);
        const bool raw_over = n_blocks > ctx->blk_cap;
        const bool cand_over = raw_over, hit_over = nh > s.hit_cap;
        if (!cand_over && !hit_over) {
            s.timing.n_candidates = n_cand; s.timing.n_hits = nh;
            s.timing.engine_used = (ctx->engine == B200SCAN_ENGINE_GATHER || !ctx->tc_usable) ? B200SCAN_ENGINE_GATHER : B200SCAN_ENGINE_TENSOR;
            (void)has_zero;
            break;
        }
        if (attempt >= 4) return fail(ctx, B200SCAN_ECUDA, "hit buffers still too small after regrowing");
        // the counters kept counting, so they say exactly how much room a re-run needs
        CU(cudaStreamSynchronize(ctx->stream));
        {
            const unsigned long long need_hits = hit_over ? nh + nh / 8 + 1024 : s.hit_cap;
            const unsigned long long need_cand = raw_over ? ((unsigned long long)n_blocks + n_blocks / 8 + 1024) * kRawBlock : ctx->cand_cap;
            if (need_hits > ctx->hit_budget || need_cand > 4 * ctx->hit_budget) {
                s.resident = false;
                return fail(ctx, B200SCAN_ENOMEM, "block too dense for the device buffers: %llu hits, %llu candidates against a budget of %llu hit records "
                            "-- submit it in smaller blocks", nh, n_cand, ctx->hit_budget);
            }
        }
        // grow with a way back: if the larger allocation fails the old size is restored, the block is dropped and the caller told
        auto regrow = [&](void** ptr, size_t old_bytes, size_t new_bytes) -> bool {
            cudaFree(*ptr); *ptr = nullptr;
            if (cudaMalloc(ptr, new_bytes) == cudaSuccess) return true;
            cudaGetLastError();
            *ptr = nullptr;
            if (cudaMalloc(ptr, old_bytes) != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; }
            return false;
        };
        bool ok = true;
        if (raw_over) {
            const uint32_t new_cap = n_blocks + n_blocks / 8 + 1024;
            const size_t raw_old = ((size_t)ctx->blk_cap + 1) * kRawBlock * kRawWords * 4, raw_new = ((size_t)new_cap + 1) * kRawBlock * kRawWords * 4;
            if (regrow(reinterpret_cast<void**>(&ctx->d_raw), raw_old, raw_new) &&
                regrow(reinterpret_cast<void**>(&ctx->d_blk_count), (size_t)ctx->blk_cap * 4, (size_t)new_cap * 4) &&
                regrow(reinterpret_cast<void**>(&ctx->d_blk_tag), (size_t)ctx->blk_cap * 4, (size_t)new_cap * 4)) {
                ctx->blk_cap = new_cap;
                ctx->cand_cap = std::max<unsigned long long>(ctx->cand_cap, (unsigned long long)ctx->blk_cap * kRawBlock);
            } else ok = false;
        }
        if (ok && (hit_over || cand_over)) {
            unsigned long long want = std::max<unsigned long long>(nh + nh / 8 + 1024, s.hit_cap);
            if (want > s.hit_cap) {
                if (regrow(reinterpret_cast<void**>(&s.d_hits), sizeof(b200scan_hit) * s.hit_cap, sizeof(b200scan_hit) * want)) s.hit_cap = want; else ok = false;
            }
        }
        if (!ok) {
            s.resident = false;
            if (!ctx->d_raw || !ctx->d_blk_count || !ctx->d_blk_tag || !s.d_hits)
                return fail(ctx, B200SCAN_ECUDA, "out of device memory while regrowing the hit buffers, and the previous size could not be restored");
            return fail(ctx, B200SCAN_ENOMEM, "out of device memory for a block with %llu hits and %llu candidates -- submit it in smaller blocks", nh, n_cand);
        }
        int rc = reset_counters(ctx, s, true, ctx->stream);
        if (rc) return rc;
        int launches = 0;
        rc = launch_scoring(ctx, s, s.ev[3], s.ev[4], &launches, s.ev[9]);
        if (rc) return rc;
        s.timing.kernel_launches += launches;
        CU(cudaMemcpyAsync(s.h_counters, s.d_counters, 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaEventRecord(s.ev[5], ctx->stream));
    }
    const unsigned long long nh = s.h_counters[1];
    {
        const int rc = ensure_host_hits(ctx, s, (size_t)s.hit_bytes * nh);
        if (rc) return rc;
    }
    const uint64_t nbk = (s.n_payload + kBucketSize - 1) >> kBucketShift;
    if (s.hit_bytes == B200SCAN_HITS_8 && s.h_bucket_cap < nbk + 1) {
        hfree(s.h_bucket);
        s.h_bucket_cap = 0;
        CU(cudaMallocHost(&s.h_bucket, (nbk + 1 + nbk / 4) * 4));
        s.h_bucket_cap = nbk + 1 + nbk / 4;
    }
    // the scan of this slot is complete (ev[5] was waited for): download its hits on the copy stream, so that the
    // kernels of a block already submitted on the other slot keep the compute stream busy meanwhile
    CU(cudaEventRecord(s.ev[6], ctx->copy_stream));
    if (nh) CU(cudaMemcpyAsync(s.h_hits, s.d_hits, (size_t)s.hit_bytes * nh, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (s.hit_bytes == B200SCAN_HITS_8) CU(cudaMemcpyAsync(s.h_bucket, s.d_bucket_start, (nbk + 1) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaEventRecord(s.ev[7], ctx->copy_stream));
    CU(cudaEventSynchronize(s.ev[7]));
    cudaEventElapsedTime(&s.timing.h2d_ms, s.ev[0], s.ev[1]);
    cudaEventElapsedTime(&s.timing.pack_ms, s.ev[1], s.ev[2]);
    cudaEventElapsedTime(&s.timing.score_ms, s.ev[8], s.ev[3]);
    cudaEventElapsedTime(&s.timing.rescore_ms, s.ev[3], s.ev[4]);
    cudaEventElapsedTime(&s.timing.d2h_ms, s.ev[6], s.ev[7]);
    cudaEventElapsedTime(&s.timing.order_ms, s.ev[4], s.ev[9]);
    if (bucket_start) *bucket_start = s.h_bucket;
    if (n_buckets) *n_buckets = nbk;
    if (hits) *hits = s.h_hits;
    if (n_hits) *n_hits = nh;
    if (timing) *timing = s.timing;
    return B200SCAN_OK;
}

int b200scan_collect(b200scan_ctx* ctx, int slot, const b200scan_hit** hits, uint64_t* n_hits, b200scan_timing* timing)
{
    const void* p = nullptr;
    const int rc = collect_impl(ctx, slot, B200SCAN_HITS_16, &p, n_hits, timing);
    if (rc == B200SCAN_OK && hits) *hits = static_cast<const b200scan_hit*>(p);
    return rc;
}

int b200scan_collect12(b200scan_ctx* ctx, int slot, const b200scan_hit12** hits, uint64_t* n_hits, b200scan_timing* timing)
{
    const void* p = nullptr;
    const int rc = collect_impl(ctx, slot, B200SCAN_HITS_12, &p, n_hits, timing);
    if (rc == B200SCAN_OK && hits) *hits = static_cast<const b200scan_hit12*>(p);
    return rc;
}

int b200scan_collect8(b200scan_ctx* ctx, int slot, const b200scan_hit8** hits, uint64_t* n_hits, const uint32_t** bucket_start,
                      uint64_t* n_buckets, b200scan_timing* timing)
{
    const void* p = nullptr;
    const int rc = collect_impl(ctx, slot, B200SCAN_HITS_8, &p, n_hits, timing, bucket_start, n_buckets);
    if (rc == B200SCAN_OK && hits) *hits = static_cast<const b200scan_hit8*>(p);
    return rc;
}

int b200scan_set_hit_format(b200scan_ctx* ctx, int format)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (format != B200SCAN_HITS_16 && format != B200SCAN_HITS_12 && format != B200SCAN_HITS_8)
        return fail(ctx, B200SCAN_EINVAL, "hit format must be B200SCAN_HITS_16, B200SCAN_HITS_12 or B200SCAN_HITS_8");
    if (format == B200SCAN_HITS_8 && ctx->have_motifs && ctx->n_cols > (1u << 24))
        return fail(ctx, B200SCAN_ELIMIT, "B200SCAN_HITS_8 records hold 24-bit column indices (%u columns loaded)", ctx->n_cols);
    for (auto& s : ctx->slot) if (s.in_flight) return fail(ctx, B200SCAN_ESTATE, "set_hit_format while a block is in flight");
    ctx->hit_format = format;
    return B200SCAN_OK;
}

int b200scan_rerun_resident(b200scan_ctx* ctx, int slot, int iters, float* total_ms, float* score_kernel_ms, uint64_t* n_hits_last)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (slot < 0 || slot >= B200SCAN_NUM_SLOTS || iters < 1 || iters > 4096) return fail(ctx, B200SCAN_EINVAL, "bad slot / iters");
    Slot& s = ctx->slot[slot];
    if (s.in_flight || !s.resident) return fail(ctx, B200SCAN_ESTATE, "slot %d holds no collected resident block", slot);
    CU(cudaSetDevice(ctx->device));
    std::vector<cudaEvent_t> ev(3 * (size_t)iters);
    for (auto& e : ev) CU(cudaEventCreate(&e));
    int rc = B200SCAN_OK;
    for (int i = 0; i < iters && rc == B200SCAN_OK; i++) {
        rc = reset_counters(ctx, s, true, ctx->stream);
        if (rc) break;
        CU(cudaEventRecord(ev[3 * i], ctx->stream));
        rc = launch_scoring(ctx, s, ev[3 * i + 1], nullptr, nullptr, ev[3 * i + 2]);      // [.. + 2]: after rescoring and, for B200SCAN_HITS_8, the ordering kernels
    }
    if (rc == B200SCAN_OK) {
        CU(cudaMemcpyAsync(s.h_counters, s.d_counters, 32, cudaMemcpyDeviceToHost, ctx->stream));
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, B200SCAN_ECUDA, "rerun failed: %s", cudaGetErrorString(e));
    }
    float tot = 0, sc = 0;
    if (rc == B200SCAN_OK) {
        for (int i = 0; i < iters; i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ev[3 * i], ev[3 * i + 2]);
            cudaEventElapsedTime(&b, ev[3 * i], ev[3 * i + 1]);
            tot += a; sc += b;
        }
        if ((uint32_t)(s.h_counters[2] >> 32)) rc = fail(ctx, B200SCAN_ECUDA, "kernel reported error flag 0x%08x", (uint32_t)(s.h_counters[2] >> 32));
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (total_ms) *total_ms = tot;
    if (score_kernel_ms) *score_kernel_ms = sc;
    if (n_hits_last) *n_hits_last = s.h_counters[1];
    return rc;
}

#if defined(B200_TRACE) || defined(B200_PHASE)
int b200scan_debug_trace(b200scan_ctx* ctx, unsigned long long* out, int n)
{
    if (!ctx || !out) return B200SCAN_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, ctx->d_trace, std::min<size_t>((size_t)n, 4 * kTraceTiles * 4) * 8, cudaMemcpyDeviceToHost));
    return B200SCAN_OK;
}
#endif

// Diagnostic, no device needed (tests/test_fold.py): the filter weights build_motifs derives for ONE column.  w = 4 * L FP32 weights
// (ACGT per position), kind = 8 (INT8 operands), 16 or 32 (FP16 operands with FP16 / FP32 accumulators), zmode = 1 for the image
// used on blocks with zero-contribution characters.  weights[4 * L] receive the operand values as real numbers, *bias the bias
// step's value (zmode), *margin the safety margin (INT8: the worst-case overshoot), *flags bit 0 "always a candidate", bit 1
// "no window can reach the threshold".
int b200scan_debug_fold(const float* w, int32_t L, float thr, int32_t kind, int32_t zmode, double margin16_scale,
                        double* weights, double* bias, double* margin, int32_t* flags)
{
    if (!w || L < 1 || L > B200SCAN_MAX_MOTIF_LEN || (kind != 8 && kind != 16 && kind != 32) || !weights || !bias || !margin || !flags)
        return B200SCAN_EINVAL;
    std::vector<float4> wc((size_t)L);
    for (int32_t j = 0; j < L; j++) wc[(size_t)j] = make_float4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    const Folded f = kind == 8 ? fold_col_i8(wc.data(), (uint32_t)L, thr, zmode)
                   : zmode ? fold_col_z(wc.data(), (uint32_t)L, thr, kind == 16, margin16_scale)
                           : fold_col(wc.data(), (uint32_t)L, thr, kind == 16, margin16_scale);
    *flags = (f.always ? 1 : 0) | (f.never ? 2 : 0);
    *margin = f.margin;
    auto half = [](uint16_t h) { __half hh; std::memcpy(&hh, &h, 2); return (double)__half2float(hh); };
    *bias = kind == 8 ? (double)f.qbias : (zmode ? half(f.bias) : 0.0);
    for (int32_t i = 0; i < 4 * L; i++)
        weights[i] = kind == 8 ? (f.q.size() == (size_t)(4 * L) ? (double)f.q[(size_t)i] : 0.0) : (f.y.size() == (size_t)(4 * L) ? half(f.y[(size_t)i]) : 0.0);
    return B200SCAN_OK;
}

int b200scan_flush_l2(b200scan_ctx* ctx)
{
    if (!ctx) return B200SCAN_EINVAL;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = 256u << 20;
    if (!ctx->d_flush) CU(cudaMalloc(&ctx->d_flush, bytes));
    CU(cudaMemsetAsync(ctx->d_flush, 0x5a, bytes, ctx->stream));
    return B200SCAN_OK;
}

int b200scan_tensor_info(const b200scan_ctx* ctx, int32_t* accumulator_bits, double* mean_margin)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (accumulator_bits) *accumulator_bits = ctx->tc_acc_bits;
    if (mean_margin) *mean_margin = ctx->cand_inflation;
    return B200SCAN_OK;
}

int b200scan_tensor_work(const b200scan_ctx* ctx, double* mma_ops_per_window, double* algorithmic_ops_per_window)
{
    if (!ctx) return B200SCAN_EINVAL;
    // one MMA of tile t covers N = n_pad columns and K = 32 (INT8) or 16 (FP16 operands) one-hot rows of one window: 2 N K
    // operations; n_k of them per window (the bias step of masked blocks is not counted)
    double ops = 0;
    for (size_t t = 0; t < ctx->ttiles.size(); t++)
        ops += 2.0 * ctx->ttiles[t].n_pad * ctx->ttiles[t].n_k * (t < ctx->n_tiles8 ? 32.0 : 16.0);
    if (mma_ops_per_window) *mma_ops_per_window = ctx->tc_usable ? ops : 0.0;
    if (algorithmic_ops_per_window) *algorithmic_ops_per_window = 8.0 * (double)ctx->sum_len;
    return B200SCAN_OK;
}

int b200scan_describe(const b200scan_ctx* ctx, int32_t* n_cols, int32_t* max_len, int32_t* n_tiles, int32_t* sm_count, uint64_t* sum_len)
{
    if (!ctx) return B200SCAN_EINVAL;
    if (n_cols) *n_cols = (int32_t)ctx->n_cols;
    if (max_len) *max_len = (int32_t)ctx->max_len;
    if (n_tiles) *n_tiles = (int32_t)ctx->ttiles.size();
    if (sm_count) *sm_count = ctx->sm_count;
    if (sum_len) *sum_len = ctx->sum_len;
    return B200SCAN_OK;
}

} // extern "C"
