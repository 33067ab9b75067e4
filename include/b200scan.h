/*
 * b200scan.h -- C ABI of the B200-native `blamm scan` hot path (libb200scan.so).
 *
 * This is the drop-in boundary for the part of biointec/blamm that scores every window of a block of
 * filtered sequence against every motif column and extracts the occurrences.  Each entry point names the
 * reference interface it replaces (paths are relative to the reference's src/):
 *
 *   b200scan_create / _destroy     cudaSetDevice + cublasCreate + the cudaMalloc/cudaFree blocks of
 *                                  PWMScan::scanThreadCUBLAS                       pwmscan.cpp:309-357, 428-436
 *   b200scan_set_motifs            upload of matrix P and the per-column thresholds  pwmscan.cpp:316-342
 *                                  (P as produced by MotifContainer::generateMatrix  motif.cpp:542-564,
 *                                   thresholds as set in the species loop            pwmscan.cpp:599-616)
 *   b200scan_submit_ascii          SeqMatrix::getNextSeqMatrix one-hot fill + cublasSetVector(S)
 *                                                                            sequence.cpp:299-340, pwmscan.cpp:383
 *                                  + the w-iteration loop { Matrix::sgemm_batch_cuda; kernel_wrapper }
 *                                                                  pwmscan.cpp:385-391, matrix.h:314-323, kernel.cu:21-45
 *   b200scan_submit_packed         same, for callers that already hold the 2-bit packed stream
 *   b200scan_collect               cublasGetVector(d_nOcc / occIdx / occScore) + the boundary filter of
 *                                  PWMScan::extractOccurrences2                      pwmscan.cpp:393-406, 135-163
 *   b200scan_collect8              same, with the block's occurrences already in (position, column) order -- what the
 *                                  reference leaves to the order of its R sweep      pwmscan.cpp:108-131 (README.md:165)
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * B200SCAN_E* code (no exception crosses the ABI); b200scan_last_error() gives the text.  The context owns
 * all device and pinned memory; inputs are copied before the call returns unless stated; outputs are
 * borrowed until the next submit on the same slot.  A context is bound to one CUDA device and must be
 * driven by one host thread at a time (the reference uses one host thread per device, pwmscan.cpp:444-447).
 * There is NO CPU fallback: without a usable sm_100 device b200scan_create fails.
 *
 * A "block" is what the reference calls SeqBlock (sequence.h:95-184): the concatenation of valid
 * (ACGTacgt) fragments of one species group, `n_total` = payload + halo characters, where only windows
 * starting in the first `n_payload` characters are reported (the halo of maxLen-1 characters belongs to the
 * next block, sequence.cpp:274-293).  `frag_starts` are the block positions at which a new fragment begins
 * (the keys of SeqBlock::block2seq, position 0 is implied); a window is an occurrence only if it lies wholly
 * inside one fragment (pwmscan.cpp:122-126).  Unlike the reference's h*w = 250,000 character blocks, a block
 * here may hold up to `max_block_nt` characters (hundreds of Mnt).
 */
#ifndef B200SCAN_H
#define B200SCAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SCAN_ABI_VERSION 2   /* 2: three slots, B200SCAN_HITS_8 / b200scan_collect8, order_ms in b200scan_timing, lazy buffers */

/* error codes */
#define B200SCAN_OK            0
#define B200SCAN_EINVAL       -1   /* bad argument */
#define B200SCAN_ECUDA        -2   /* CUDA runtime error (text in last_error) */
#define B200SCAN_ENODEVICE    -3   /* no sm_100 device / device index out of range */
#define B200SCAN_ENOMEM       -4   /* host or device allocation failed; from b200scan_collect*: block too dense for the hit budget (see b200scan_create) */
#define B200SCAN_ESTATE       -5   /* call order violated (e.g. collect without submit) */
#define B200SCAN_ELIMIT       -6   /* motif longer than B200SCAN_MAX_MOTIF_LEN, block larger than max_block_nt */

#define B200SCAN_MAX_MOTIF_LEN 64
#define B200SCAN_NUM_SLOTS      3  /* blocks in flight per context: one uploading, one being scored, one whose hits are downloading.
                                    * A slot's device buffers are allocated at its first submit, so a caller that alternates two
                                    * slots (the CLI) pays for two. */
#define B200SCAN_BUCKET_SHIFT   8  /* B200SCAN_HITS_8: hits are grouped in buckets of 2^8 = 256 window positions */

/* scoring engines (b200scan_set_engine) */
#define B200SCAN_ENGINE_AUTO    0  /* tensor-core filter + exact rescore (gather-add only if the motif set is unusable for it) */
#define B200SCAN_ENGINE_GATHER  1  /* shared-memory gather-add, every score in exact reference order */
#define B200SCAN_ENGINE_TENSOR  2  /* tcgen05 filter + exact rescore, also for blocks with zero-contribution characters   */

/* how lower-case acgt is scored (b200scan_submit_ascii) */
#define B200SCAN_LOWER_ZERO     0  /* BLAS-path semantics: valid character, contributes 0 (sequence.cpp:312-319) */
#define B200SCAN_LOWER_FOLD     1  /* naive-path semantics: scored like upper case (motif.cpp:138-149)           */

typedef struct b200scan_ctx b200scan_ctx;

/* One occurrence.  `pos` is the block position of the window start (0-based, < n_payload), `col` the motif
 * column as passed to b200scan_set_motifs, `score` the FP32 log-odds score summed in position order
 * (the naive in-order sum: bit-identical to the reference's naive path, motif.cpp:225-239; its BLAS path re-associates the sums of
 * longer motifs and differs from that by a few ulp -- an occurrence whose score lies that close to its threshold can differ between
 * the two; north_star's tolerance for such cases is 1e-4, tools/parity_list.py lists them). */
typedef struct b200scan_hit {
    uint64_t pos;
    uint32_t col;
    float    score;
} b200scan_hit;

/* The same occurrence as a 12-byte record: block positions always fit 32 bits (max_block_nt < 2^32).  A quarter less
 * device->host traffic -- the hit list IS the PCIe traffic of this path (4 bytes of hits come back for every byte of
 * sequence sent at -pt 1e-4 with 1800 columns), and on an 8-GPU box the ranks share the host link.  Selected per
 * context with b200scan_set_hit_format and fetched with b200scan_collect12. */
typedef struct b200scan_hit12 {
    uint32_t pos;
    uint32_t col;
    float    score;
} b200scan_hit12;
#define B200SCAN_HITS_16 16   /* b200scan_hit records, b200scan_collect (default) */
#define B200SCAN_HITS_12 12   /* b200scan_hit12 records, b200scan_collect12        */
#define B200SCAN_HITS_8   8   /* b200scan_hit8 records IN (position, column) ORDER, b200scan_collect8 */

/* The same occurrence as an 8-byte record of an ORDERED list (B200SCAN_HITS_8): the device sorts the block's hits by
 * (position, column) (csrc/order.cuh) and groups them in buckets of 256 window positions; b200scan_collect8 returns the
 * records together with bucket_start[0 .. n_buckets], n_buckets = ceil(n_payload / 256): the hits of bucket b are records
 * bucket_start[b] .. bucket_start[b + 1] - 1 and record i of them lies at block position 256 b + (key >> 24); its column is
 * key & 0xFFFFFF.  A third less device->host traffic than b200scan_hit12 (plus 4 bytes per 256 positions), and the host no
 * longer sorts.  Column indices must fit 24 bits (n_cols <= 2^24: b200scan_set_hit_format checks at the next set_motifs). */
typedef struct b200scan_hit8 {
    uint32_t key;      /* (pos & 255) << 24 | col */
    float    score;
} b200scan_hit8;

/* timings of the last launch on a slot, CUDA events on the context's stream (milliseconds) */
typedef struct b200scan_timing {
    float h2d_ms;        /* host->device copy of the block (+ fragment table)          */
    float pack_ms;       /* ASCII -> 2-bit pack kernel                                 */
    float score_ms;      /* dominant kernel: tensor filter or gather-add               */
    float rescore_ms;    /* exact rescore + boundary filter + compaction (tensor path) */
    float d2h_ms;        /* hit download                                               */
    uint64_t n_candidates;   /* tensor path: candidates the filter passed to the rescorer */
    uint64_t n_hits;
    int32_t  engine_used;    /* B200SCAN_ENGINE_GATHER or _TENSOR */
    int32_t  kernel_launches;/* kernels launched for this block */
    float order_ms;      /* B200SCAN_HITS_8: bucket scan + scatter + order kernels (0 otherwise)   */
    float reserved;
} b200scan_timing;

int  b200scan_abi_version(void);

/* Number of usable (sm_100) CUDA devices, 0 if none.  Replaces cudaGetDeviceCount in PWMScan's ctor (pwmscan.cpp:563). */
int  b200scan_device_count(void);

/* Create a context on CUDA device `device`.  max_block_nt: largest n_total a submit may carry.
 * max_hits: initial capacity (records) of the per-slot hit buffers.  A block that produces more hits (or more filter
 * candidates) than the buffers hold is NOT lost: every counter keeps counting past its capacity, b200scan_collect then
 * frees the buffers, allocates them at the counted size (+ 1/8) and scores the whole block again -- at the price of that
 * second pass and of device memory of about 96 bytes per hit of the densest block.  The growth has a ceiling: a budget of
 * hit records derived from the free device memory at creation (60 % of it; B200SCAN_HIT_BUDGET=<records> overrides).  A block
 * that needs more, or whose larger allocation fails (the previous size is then restored), makes b200scan_collect* return
 * B200SCAN_ENOMEM: the slot is free again, nothing was returned for the block, and the caller submits it in smaller pieces --
 * e.g. its two halves, each with the maxLen - 1 characters behind it as halo (what the CLI does, cli.cpp: scanSplit).
 * Device and pinned buffers of a slot are allocated at the slot's first use; the pinned hit buffer is sized from the
 * blocks actually collected. */
int  b200scan_create(b200scan_ctx** out, int device, uint64_t max_block_nt, uint64_t max_hits);
void b200scan_destroy(b200scan_ctx* ctx);
const char* b200scan_last_error(const b200scan_ctx* ctx);   /* ctx may be NULL: error of the last failed create */

int  b200scan_set_engine(b200scan_ctx* ctx, int engine);

/* Operand / accumulator kind of the tensor-core filter: 0 = automatic (per column tile: INT8 operands with exact S32
 * accumulation where the integer weights resolve every column of the tile; else FP16 operands with FP16 accumulators in
 * TMEM when the error bound computed in b200scan_set_motifs allows it, else FP32 accumulators); 8, 16 or 32 to force
 * one kind everywhere.  Takes effect at the next b200scan_set_motifs.  Results do not depend on it (the filter is
 * conservative, scores are re-summed exactly). */
int  b200scan_set_tensor_accumulator(b200scan_ctx* ctx, int bits);
/* What the last b200scan_set_motifs chose (8, 16, 32, or 0 when the tiles differ), and the mean safety margin (score
 * units) folded into the filter (INT8: the mean worst-case overshoot of the integer weights). */
int  b200scan_tensor_info(const b200scan_ctx* ctx, int32_t* accumulator_bits, double* mean_margin);

/* Motif matrix of the current species: P is column-major with leading dimension ldp >= 4*max(col_len),
 * P[col*ldp + 4*j + o] = PWM_col[j][o], o in ACGT order (motif.cpp:556-563); col_len[col] in
 * [1, B200SCAN_MAX_MOTIF_LEN]; thr[col] the score cut-off (hit <=> score >= thr, pwmscan.cpp:115-117). */
int  b200scan_set_motifs(b200scan_ctx* ctx, const float* P, int32_t ldp, int32_t n_cols,
                         const int32_t* col_len, const float* thr);

/* Page-locked host memory for blocks (cudaMallocHost / cudaFreeHost).  A block submitted from such memory is
 * read by the copy engine directly and must stay unchanged until b200scan_collect; a block in ordinary
 * pageable memory is copied to the slot's staging buffer before b200scan_submit_ascii returns. */
void* b200scan_host_alloc(uint64_t bytes);
void  b200scan_host_free(void* p);

/* Asynchronous scan of one block held in host memory.  `block` = n_total characters from "ACGTacgt"
 * (anything else is scored as a zero contribution).  frag_starts: n_frag ascending block positions in
 * (0, n_total) where a new fragment starts; may be NULL when n_frag == 0 (copied before the call returns).
 * slot in [0, B200SCAN_NUM_SLOTS). */
int  b200scan_submit_ascii(b200scan_ctx* ctx, int slot, const char* block, uint64_t n_total,
                           uint64_t n_payload, const uint64_t* frag_starts, uint64_t n_frag,
                           int lowercase_mode);

/* Same for a pre-packed block: codes2 holds 2 bits per character (A=0,C=1,G=2,T=3; character i in bits
 * 2*(i%16) of 32-bit word i/16); zero_mask (may be NULL) holds 1 bit per character (bit i%32 of word i/32),
 * set = the character contributes 0.  Lifetime of the sources: the copy engine reads codes2 / zero_mask asynchronously when
 * they lie in page-locked memory (b200scan_host_alloc) -- such buffers must stay unchanged until the block has been
 * collected, exactly as for b200scan_submit_ascii; pageable sources are staged by the CUDA runtime before the call returns. */
int  b200scan_submit_packed(b200scan_ctx* ctx, int slot, const uint32_t* codes2, const uint32_t* zero_mask,
                            uint64_t n_total, uint64_t n_payload, const uint64_t* frag_starts,
                            uint64_t n_frag);

/* Wait for the slot's scan and return its occurrences (unordered; owned by ctx until the next submit on
 * the slot).  timing may be NULL. */
int  b200scan_collect(b200scan_ctx* ctx, int slot, const b200scan_hit** hits, uint64_t* n_hits,
                      b200scan_timing* timing);
/* Record format of the hit lists (B200SCAN_HITS_16 | B200SCAN_HITS_12 | B200SCAN_HITS_8) for blocks submitted from now on; no block may be
 * in flight.  A block must be collected with the function of the format it was submitted under (else B200SCAN_ESTATE). */
int  b200scan_set_hit_format(b200scan_ctx* ctx, int format);
int  b200scan_collect12(b200scan_ctx* ctx, int slot, const b200scan_hit12** hits, uint64_t* n_hits,
                        b200scan_timing* timing);
/* B200SCAN_HITS_8: ordered records + bucket index (see b200scan_hit8); *n_buckets = ceil(n_payload / 256), bucket_start has
 * *n_buckets + 1 entries.  Both arrays are owned by ctx until the next submit on the slot. */
int  b200scan_collect8(b200scan_ctx* ctx, int slot, const b200scan_hit8** hits, uint64_t* n_hits,
                       const uint32_t** bucket_start, uint64_t* n_buckets, b200scan_timing* timing);

/* Empirical score histograms (`blamm hist -e`, reference Histogram::histThread + extractObsScore, hist.cpp:70-140):
 * same scoring as the scan, but instead of thresholding every window that lies inside one fragment adds 1 to bin
 * clamp(int((score - col_min) / width), 0, num_bins-1) of its column, width = (col_max - col_min) / num_bins in
 * float (ScoreHistogram, motif.h:62-66, 96-102).  Call after b200scan_set_motifs (thresholds are ignored):
 *   hist_begin (zeroes the device histograms) -> hist_block_ascii for every block of the stream -> hist_read
 *   (counts[col * num_bins + bin], column order as passed to set_motifs).  Blocks are accumulated in stream order. */
int  b200scan_hist_begin(b200scan_ctx* ctx, const float* col_min, const float* col_max, uint32_t num_bins);
int  b200scan_hist_block_ascii(b200scan_ctx* ctx, const char* block, uint64_t n_total, uint64_t n_payload,
                               const uint64_t* frag_starts, uint64_t n_frag, int lowercase_mode);
int  b200scan_hist_read(b200scan_ctx* ctx, uint64_t* counts, uint64_t n_counts);

/* Measurement hook (bench.py `value`): re-run the scoring kernels `iters` times on the block that is
 * already resident in the slot's device buffers (after a submit+collect), timed with CUDA events on the
 * context's stream.  Returns total milliseconds for all iterations and for the dominant kernel alone. */
int  b200scan_rerun_resident(b200scan_ctx* ctx, int slot, int iters, float* total_ms, float* score_kernel_ms,
                             uint64_t* n_hits_last);

/* Measurement hygiene: overwrite a 256 MiB scratch buffer on the context's stream (evicts the 126 MB L2). */
int  b200scan_flush_l2(b200scan_ctx* ctx);

/* Introspection: number of window x column scores one pass over the slot's block computes. */
int  b200scan_describe(const b200scan_ctx* ctx, int32_t* n_cols, int32_t* max_len, int32_t* n_tiles,
                       int32_t* sm_count, uint64_t* sum_len);

/* Roofline bookkeeping: tensor-core operations per WINDOW of the loaded motif set -- as issued by the filter
 * (every column tile padded to its N and to whole K steps: 2 N K n_k per tile; 0 if the set runs on the gather engine) and
 * as the algorithm needs (8 x sum of the column lengths, un-padded).  What the reference spends per window is the same
 * 8 sum L inside sgemm (matrix.h:295-304). */
int  b200scan_tensor_work(const b200scan_ctx* ctx, double* mma_ops_per_window, double* algorithmic_ops_per_window);

#ifdef __cplusplus
}
#endif
#endif /* B200SCAN_H */
