// TEST INFRASTRUCTURE ONLY -- never built by `make`, never shipped, never loaded by the product.
//
// A stand-in for libb200scan.so on machines WITHOUT a GPU: the subset of include/b200scan.h that the blamm-b200 command line
// drives, answered by the CPU oracle (oracle/oracle.c: oracle_scan_stream, oracle_empirical_hist).  tests/test_cli_mock.py builds it into
// tests/mock/_build/libb200scan.so and puts that directory in front of the CLI's library search path, so that the CLI's HOST
// logic -- reader and packer, one worker per "device", group changes, three slots in flight, chunks refused as too dense and
// scored in halves, formatting, stream-order emission with several devices -- runs on the CPU suite against the reference
// binary's output.  It says nothing about the CUDA path (the -m gpu tests do that) and is not a CPU fallback: the product's
// library has none (b200scan_create fails with B200SCAN_ENODEVICE without an sm_100 device, tests/test_host.py).
//
// Knobs (environment): MOCK_B200SCAN_DEVICES (device count, default 1), MOCK_B200SCAN_HIT_BUDGET (a block with more hits
// makes collect return B200SCAN_ENOMEM, like a block too dense for the device buffers), MOCK_B200SCAN_DELAY_US (collect sleeps a
// pseudo-random time below this bound, device dependent: the devices finish out of order), MOCK_B200SCAN_SYNTH_HITS (hits per
// window and column: no scoring, a pseudo-random ordered hit list of that density -- a benchmark of the CLI's host pipeline),
// MOCK_B200SCAN_FAIL_SUBMIT / _COLLECT / _MOTIFS / _HIST = n (fault injection: the process's n-th such call fails with B200SCAN_ECUDA).
#include "../../include/b200scan.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <string>
#include <thread>
#include <vector>

extern "C" uint64_t oracle_scan_stream(const char* stream, uint64_t n, uint64_t n_payload, const uint64_t* frag_start, uint64_t n_frag,
                                       const float* P, int ldp, int n_cols, const int32_t* col_len, const float* thr, int lower_fold,
                                       uint64_t* hit_pos, uint32_t* hit_col, float* hit_score, uint64_t cap);
extern "C" int oracle_empirical_hist(const char* stream, uint64_t n, const uint64_t* frag_start, uint64_t n_frag, const float* P, int ldp, int n_cols,
                                     const int32_t* col_len, const float* mins, const float* maxs, int num_bins, int lower_fold, uint64_t* counts);

namespace {
struct Slot {
    bool in_flight = false; int fmt = B200SCAN_HITS_16; bool too_dense = false;
    std::vector<b200scan_hit> h16; std::vector<b200scan_hit12> h12; std::vector<b200scan_hit8> h8; std::vector<uint32_t> buckets;
    uint64_t n_payload = 0, n_hits = 0;
    std::future<void> pending;                  // synthetic-hit mode: the list is made in the background, like a kernel on a device
};
thread_local std::string g_create_error;
// fault injection: the n-th submit / collect / set_motifs of the process (counted over all devices, from 1) fails with B200SCAN_ECUDA
std::atomic<long> g_submits{0}, g_collects{0}, g_motifs{0}, g_hist_blocks{0};
bool inject(const char* var, std::atomic<long>& counter)
{
    const char* e = getenv(var);
    const long n = ++counter;
    return e && atol(e) == n;
}
}

struct b200scan_ctx {
    int device = 0; std::string err; int fmt = B200SCAN_HITS_16;
    std::vector<float> P; int ldp = 0, n_cols = 0; std::vector<int32_t> len; std::vector<float> thr;
    Slot slot[B200SCAN_NUM_SLOTS];
    uint64_t max_block = 0, budget = ~0ull; unsigned delay_us = 0; uint64_t rng = 1; double synth_rate = 0;
    std::vector<float> hmin, hmax; uint32_t hbins = 0; std::vector<uint64_t> hist;
};

namespace {
int fail(b200scan_ctx* c, int code, const char* fmt, ...)
{
    char buf[256]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

int scan_block(b200scan_ctx* c, int slot, const std::string& chars, uint64_t n_payload, const uint64_t* frag, uint64_t n_frag, int lower_fold)
{
    if (!c) return B200SCAN_EINVAL;
    if (slot < 0 || slot >= B200SCAN_NUM_SLOTS) return fail(c, B200SCAN_EINVAL, "slot %d out of range", slot);
    if (c->n_cols == 0) return fail(c, B200SCAN_ESTATE, "no motifs loaded");
    if (chars.size() > c->max_block) return fail(c, B200SCAN_ELIMIT, "block larger than max_block_nt");
    if (n_payload > chars.size()) return fail(c, B200SCAN_EINVAL, "n_payload > n_total");
    Slot& s = c->slot[slot];
    if (s.in_flight) return fail(c, B200SCAN_ESTATE, "slot %d is in flight", slot);
    if (inject("MOCK_B200SCAN_FAIL_SUBMIT", g_submits)) return fail(c, B200SCAN_ECUDA, "injected failure of a submit (mock)");
    std::vector<uint64_t> fs(1, 0);
    for (uint64_t i = 0; i < n_frag; i++) {
        if (frag[i] == 0 || frag[i] >= chars.size() || frag[i] <= fs.back()) return fail(c, B200SCAN_EINVAL, "fragment starts must ascend inside (0, n_total)");
        fs.push_back(frag[i]);
    }
    std::vector<uint64_t> pos(1 << 16); std::vector<uint32_t> col(1 << 16); std::vector<float> sc(1 << 16);
    uint64_t nh = 0;
    if (c->synth_rate > 0 && c->fmt == B200SCAN_HITS_8) {
        // host-pipeline benchmark mode (tools/host_pipeline_bench.sh): no scoring at all, a pseudo-random hit list of the given
        // density per (window, column) in (position, column) order -- what the reader, formatter and writer have to keep up with.
        // Written straight into the ordered 8-byte records: the stand-in should cost as little as a GPU does.
        const double per_pos = c->synth_rate * c->n_cols;                  // expected hits per window position
        const uint64_t n_total = chars.size(), n_cols = (uint64_t)c->n_cols;
        s.in_flight = true; s.fmt = c->fmt; s.n_payload = n_payload; s.too_dense = false;
        s.pending = std::async(std::launch::async, [&s, per_pos, n_total, n_payload, n_cols] {
            const uint64_t nb = (n_payload + (1u << B200SCAN_BUCKET_SHIFT) - 1) >> B200SCAN_BUCKET_SHIFT;
            s.h16.clear(); s.h12.clear(); s.h8.clear(); s.buckets.assign(nb + 1, 0);
            s.h8.reserve((size_t)(per_pos * (double)n_payload * 1.05) + 1024);
            uint64_t x = 0x9E3779B97F4A7C15ull ^ (n_total * 1315423911ull);
            const double step = 1.0 / 1000.0 / per_pos;
            double p = 0; uint64_t last = ~0ull;
            for (;;) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                p += (double)(1 + (x >> 20) % 2000) * step;                  // mean gap 1 / per_pos
                if (p >= (double)n_payload) break;
                const uint64_t q = (uint64_t)p;
                if (q == last) continue;                                      // one hit per position keeps the (position, column) order trivially
                last = q;
                s.h8.push_back({(uint32_t)((q & 255u) << 24) | (uint32_t)((x >> 3) % n_cols), 5.0f + (float)(x % 100000) / 1e4f});
                s.buckets[(q >> B200SCAN_BUCKET_SHIFT) + 1]++;
            }
            for (uint64_t b = 0; b < nb; b++) s.buckets[b + 1] += s.buckets[b];
            s.n_hits = s.h8.size();
        });
        return B200SCAN_OK;
    } else
    for (int attempt = 0; attempt < 2; attempt++) {
        nh = oracle_scan_stream(chars.data(), chars.size(), n_payload, fs.data(), fs.size(), c->P.data(), c->ldp, c->n_cols, c->len.data(),
                                c->thr.data(), lower_fold, pos.data(), col.data(), sc.data(), pos.size());
        if (nh <= pos.size()) break;
        pos.resize(nh); col.resize(nh); sc.resize(nh);
    }
    s.in_flight = true; s.fmt = c->fmt; s.n_payload = n_payload; s.n_hits = nh; s.too_dense = nh > c->budget;
    s.h16.clear(); s.h12.clear(); s.h8.clear(); s.buckets.clear();
    if (s.too_dense) return B200SCAN_OK;
    // the oracle lists the hits by position, then column: the order B200SCAN_HITS_8 promises
    if (s.fmt == B200SCAN_HITS_8) {
        const uint64_t nb = (n_payload + (1u << B200SCAN_BUCKET_SHIFT) - 1) >> B200SCAN_BUCKET_SHIFT;
        s.buckets.assign(nb + 1, 0);
        for (uint64_t i = 0; i < nh; i++) {
            s.h8.push_back({(uint32_t)((pos[i] & 255u) << 24) | col[i], sc[i]});
            s.buckets[(pos[i] >> B200SCAN_BUCKET_SHIFT) + 1]++;
        }
        for (uint64_t b = 0; b < nb; b++) s.buckets[b + 1] += s.buckets[b];
    } else {
        // the unordered formats: hand the list over back to front, so that a caller that forgets to sort is found out
        for (uint64_t i = nh; i-- > 0;) {
            if (s.fmt == B200SCAN_HITS_12) s.h12.push_back({(uint32_t)pos[i], col[i], sc[i]});
            else s.h16.push_back({pos[i], col[i], sc[i]});
        }
    }
    return B200SCAN_OK;
}

int collect_common(b200scan_ctx* c, int slot, int fmt, b200scan_timing* timing)
{
    if (!c) return B200SCAN_EINVAL;
    if (slot < 0 || slot >= B200SCAN_NUM_SLOTS) return fail(c, B200SCAN_EINVAL, "slot %d out of range", slot);
    Slot& s = c->slot[slot];
    if (!s.in_flight) return fail(c, B200SCAN_ESTATE, "slot %d has nothing to collect", slot);
    if (s.fmt != fmt) return fail(c, B200SCAN_ESTATE, "slot %d was submitted with %d-byte hit records", slot, s.fmt);
    s.in_flight = false;
    if (s.pending.valid()) s.pending.get();
    if (inject("MOCK_B200SCAN_FAIL_COLLECT", g_collects)) return fail(c, B200SCAN_ECUDA, "injected failure of a collect (mock)");
    if (c->delay_us) {
        c->rng = c->rng * 6364136223846793005ull + 1442695040888963407ull;
        std::this_thread::sleep_for(std::chrono::microseconds((c->rng >> 33) % c->delay_us));
    }
    if (s.too_dense) return fail(c, B200SCAN_ENOMEM, "block too dense for the device buffers: %llu hits against a budget of %llu (mock)",
                                 (unsigned long long)s.n_hits, (unsigned long long)c->budget);
    if (timing) { std::memset(timing, 0, sizeof *timing); timing->n_hits = s.n_hits; timing->n_candidates = s.n_hits; timing->engine_used = B200SCAN_ENGINE_GATHER; }
    return B200SCAN_OK;
}
}

extern "C" {

int b200scan_abi_version(void) { return B200SCAN_ABI_VERSION; }
int b200scan_device_count(void) { const char* e = getenv("MOCK_B200SCAN_DEVICES"); return e ? std::max(0, atoi(e)) : 1; }
const char* b200scan_last_error(const b200scan_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int b200scan_create(b200scan_ctx** out, int device, uint64_t max_block_nt, uint64_t)
{
    if (!out) return B200SCAN_EINVAL;
    *out = nullptr;
    if (device < 0 || device >= b200scan_device_count()) return fail(nullptr, B200SCAN_ENODEVICE, "device %d out of range (mock)", device);
    b200scan_ctx* c = new b200scan_ctx;
    c->device = device; c->max_block = max_block_nt; c->rng = 77 + (uint64_t)device;
    if (const char* e = getenv("MOCK_B200SCAN_HIT_BUDGET")) c->budget = strtoull(e, nullptr, 10);
    if (const char* e = getenv("MOCK_B200SCAN_DELAY_US")) c->delay_us = (unsigned)atoi(e);
    if (const char* e = getenv("MOCK_B200SCAN_SYNTH_HITS")) c->synth_rate = atof(e);
    *out = c;
    return B200SCAN_OK;
}
void b200scan_destroy(b200scan_ctx* c) { delete c; }
int b200scan_set_engine(b200scan_ctx* c, int engine) { return c && engine >= 0 && engine <= 2 ? B200SCAN_OK : B200SCAN_EINVAL; }
int b200scan_set_tensor_accumulator(b200scan_ctx* c, int) { return c ? B200SCAN_OK : B200SCAN_EINVAL; }

int b200scan_set_hit_format(b200scan_ctx* c, int format)
{
    if (!c) return B200SCAN_EINVAL;
    if (format != B200SCAN_HITS_16 && format != B200SCAN_HITS_12 && format != B200SCAN_HITS_8) return fail(c, B200SCAN_EINVAL, "bad hit format");
    for (const Slot& s : c->slot) if (s.in_flight) return fail(c, B200SCAN_ESTATE, "set_hit_format while a block is in flight");
    c->fmt = format;
    return B200SCAN_OK;
}

int b200scan_set_motifs(b200scan_ctx* c, const float* P, int32_t ldp, int32_t n_cols, const int32_t* col_len, const float* thr)
{
    if (!c || !P || !col_len || !thr || n_cols < 1) return B200SCAN_EINVAL;
    for (const Slot& s : c->slot) if (s.in_flight) return fail(c, B200SCAN_ESTATE, "set_motifs while a block is in flight");
    if (inject("MOCK_B200SCAN_FAIL_MOTIFS", g_motifs)) return fail(c, B200SCAN_ECUDA, "injected failure of set_motifs (mock)");
    for (int32_t i = 0; i < n_cols; i++)
        if (col_len[i] < 1 || col_len[i] > B200SCAN_MAX_MOTIF_LEN || 4 * col_len[i] > ldp) return fail(c, B200SCAN_ELIMIT, "column %d: bad length", i);
    c->P.assign(P, P + (size_t)ldp * n_cols); c->ldp = ldp; c->n_cols = n_cols;
    c->len.assign(col_len, col_len + n_cols); c->thr.assign(thr, thr + n_cols);
    return B200SCAN_OK;
}

int b200scan_submit_ascii(b200scan_ctx* c, int slot, const char* block, uint64_t n_total, uint64_t n_payload, const uint64_t* frag_starts,
                          uint64_t n_frag, int lowercase_mode)
{
    if (!c || (!block && n_total)) return B200SCAN_EINVAL;
    return scan_block(c, slot, std::string(block, block + n_total), n_payload, frag_starts, n_frag, lowercase_mode == B200SCAN_LOWER_FOLD);
}

int b200scan_submit_packed(b200scan_ctx* c, int slot, const uint32_t* codes2, const uint32_t* zero_mask, uint64_t n_total, uint64_t n_payload,
                           const uint64_t* frag_starts, uint64_t n_frag)
{
    if (!c || (!codes2 && n_total)) return B200SCAN_EINVAL;
    std::string chars(n_total, 'A');
    for (uint64_t i = 0; i < n_total && c->synth_rate <= 0; i++) {
        const char up = "ACGT"[(codes2[i >> 4] >> (2 * (i & 15))) & 3u];
        const bool zero = zero_mask && ((zero_mask[i >> 5] >> (i & 31)) & 1u);
        chars[i] = zero ? (char)(up | 0x20) : up;             // a zero-contribution character = lower case under the BLAS-path rule
    }
    return scan_block(c, slot, chars, n_payload, frag_starts, n_frag, 0);
}

int b200scan_collect(b200scan_ctx* c, int slot, const b200scan_hit** hits, uint64_t* n_hits, b200scan_timing* timing)
{
    const int rc = collect_common(c, slot, B200SCAN_HITS_16, timing);
    if (rc) return rc;
    *hits = c->slot[slot].h16.data(); *n_hits = c->slot[slot].h16.size();
    return B200SCAN_OK;
}
int b200scan_collect12(b200scan_ctx* c, int slot, const b200scan_hit12** hits, uint64_t* n_hits, b200scan_timing* timing)
{
    const int rc = collect_common(c, slot, B200SCAN_HITS_12, timing);
    if (rc) return rc;
    *hits = c->slot[slot].h12.data(); *n_hits = c->slot[slot].h12.size();
    return B200SCAN_OK;
}
int b200scan_collect8(b200scan_ctx* c, int slot, const b200scan_hit8** hits, uint64_t* n_hits, const uint32_t** bucket_start, uint64_t* n_buckets,
                      b200scan_timing* timing)
{
    const int rc = collect_common(c, slot, B200SCAN_HITS_8, timing);
    if (rc) return rc;
    Slot& s = c->slot[slot];
    *hits = s.h8.data(); *n_hits = s.h8.size(); *bucket_start = s.buckets.data(); *n_buckets = s.buckets.size() - 1;
    return B200SCAN_OK;
}

// `hist -e`: every window that STARTS in the payload and lies inside one fragment, binned as the oracle bins it.  The oracle counts
// every window of a stream, so a block's share is  counts(whole block) - counts(halo alone)  (a window starting in the halo that
// fits the block fits the halo, and the fragment rule looks only forward).
int b200scan_hist_begin(b200scan_ctx* c, const float* col_min, const float* col_max, uint32_t num_bins)
{
    if (!c || !col_min || !col_max || num_bins < 1) return B200SCAN_EINVAL;
    if (c->n_cols == 0) return fail(c, B200SCAN_ESTATE, "no motifs loaded");
    c->hmin.assign(col_min, col_min + c->n_cols); c->hmax.assign(col_max, col_max + c->n_cols);
    c->hbins = num_bins; c->hist.assign((size_t)c->n_cols * num_bins, 0);
    return B200SCAN_OK;
}
int b200scan_hist_block_ascii(b200scan_ctx* c, const char* block, uint64_t n_total, uint64_t n_payload, const uint64_t* frag_starts, uint64_t n_frag,
                              int lowercase_mode)
{
    if (!c || (!block && n_total)) return B200SCAN_EINVAL;
    if (!c->hbins) return fail(c, B200SCAN_ESTATE, "b200scan_hist_begin has not been called");
    if (n_total > c->max_block || n_payload > n_total) return fail(c, B200SCAN_ELIMIT, "block larger than max_block_nt");
    if (inject("MOCK_B200SCAN_FAIL_HIST", g_hist_blocks)) return fail(c, B200SCAN_ECUDA, "injected failure of a histogram block (mock)");
    std::vector<uint64_t> fs(1, 0), fs_halo(1, 0);
    for (uint64_t i = 0; i < n_frag; i++) {
        fs.push_back(frag_starts[i]);
        if (frag_starts[i] > n_payload) fs_halo.push_back(frag_starts[i] - n_payload);
    }
    std::vector<uint64_t> all((size_t)c->n_cols * c->hbins, 0), halo(all.size(), 0);
    const int fold = lowercase_mode == B200SCAN_LOWER_FOLD;
    if (oracle_empirical_hist(block, n_total, fs.data(), fs.size(), c->P.data(), c->ldp, c->n_cols, c->len.data(), c->hmin.data(), c->hmax.data(),
                              (int)c->hbins, fold, all.data()) != 0) return fail(c, B200SCAN_ENOMEM, "oracle_empirical_hist failed");
    if (n_total > n_payload &&
        oracle_empirical_hist(block + n_payload, n_total - n_payload, fs_halo.data(), fs_halo.size(), c->P.data(), c->ldp, c->n_cols, c->len.data(),
                              c->hmin.data(), c->hmax.data(), (int)c->hbins, fold, halo.data()) != 0) return fail(c, B200SCAN_ENOMEM, "oracle_empirical_hist failed");
    for (size_t i = 0; i < all.size(); i++) c->hist[i] += all[i] - halo[i];
    return B200SCAN_OK;
}
int b200scan_hist_read(b200scan_ctx* c, uint64_t* counts, uint64_t n_counts)
{
    if (!c || !counts) return B200SCAN_EINVAL;
    if (!c->hbins) return fail(c, B200SCAN_ESTATE, "b200scan_hist_begin has not been called");
    if (n_counts != c->hist.size()) return fail(c, B200SCAN_EINVAL, "hist_read: expected %llu counts", (unsigned long long)c->hist.size());
    std::copy(c->hist.begin(), c->hist.end(), counts);
    return B200SCAN_OK;
}

} // extern "C"
