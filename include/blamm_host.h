/*
 * blamm_host.h -- C ABI over the C++ host model (libblammhost.so) for callers that are not C++ (tests and
 * bench.py drive it through ctypes).  It mirrors the reference's host-side objects on the scan path:
 *
 *   blamm_motifs_*   MotifContainer: load / addReverseComplements / generateMatrix / thresholds
 *                    (motif.cpp:323-338, 439-449, 542-564; pwmscan.cpp:599-616) and the theoretical
 *                    histogram writer of `blamm hist` (hist.cpp:162-175, motif.cpp:71-107, 151-192)
 *   blamm_fasta_*    FastaBatch::getNextOverlappingBlock with its SeqBlock markers (sequence.cpp:274-293)
 *   blamm_pack_ascii / blamm_fasta_pack
 *                    the operand fill of SeqMatrix::getNextSeqMatrix (sequence.cpp:299-340), as 2-bit codes + zero mask
 *   blamm_format_score
 *                    the score column of PWMScan::writeOccToDisk (pwmscan.cpp:88-95: ostream << float)
 *
 * Unless stated otherwise functions return 0 on success, -1 on error (text from blamm_host_last_error, thread local).
 */
#ifndef BLAMM_HOST_H
#define BLAMM_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct blamm_motifs blamm_motifs;
typedef struct blamm_fasta blamm_fasta;

const char* blamm_host_last_error(void);

int  blamm_motifs_load(const char* path, int load_permutations, int add_revcompl, blamm_motifs** out);
void blamm_motifs_free(blamm_motifs* m);
int  blamm_motifs_count(const blamm_motifs* m);          /* columns (motifs x strands) */
int  blamm_motifs_max_len(const blamm_motifs* m);
/* PWMs + matrix P for background counts bg[4] (ACGT).  P is column major with ld = 4*max_len. */
int  blamm_motifs_generate_matrix(blamm_motifs* m, const uint64_t bg[4], float pseudocount);
int  blamm_motifs_get_matrix(const blamm_motifs* m, float* P, uint64_t n_floats);
int  blamm_motifs_get_columns(const blamm_motifs* m, int32_t* col_len, uint8_t* is_revcompl, float* min_score, float* max_score);
const char* blamm_motifs_name(const blamm_motifs* m, int col);
/* mode: 0 absolute (-at), 1 relative (-rt), 2 p-value (-pt, histograms hist_<species>_<motif>.dat in histdir) */
int  blamm_motifs_set_thresholds(blamm_motifs* m, int mode, float value, const char* species, const char* histdir, float* thr_out);
/* `blamm hist` (theoretical) for every forward motif of the set under background bg */
int  blamm_motifs_write_histograms(blamm_motifs* m, const uint64_t bg[4], float pseudocount, uint64_t num_bins,
                                   uint64_t max_length, const char* species, const char* histdir);

int  blamm_fasta_open(const char* const* files, int n_files, uint64_t max_filtered, blamm_fasta** out);
void blamm_fasta_close(blamm_fasta* f);
/* Parser threads (the reference reads under one mutex, sequence.cpp:274-293); segment_bytes = 0 sizes the per-thread
 * byte segments from the request, any other value forces it (tests). */
int  blamm_fasta_set_parallel(blamm_fasta* f, unsigned threads, uint64_t segment_bytes);
/* Next chunk (payload + halo).  Returns 1 if a chunk was produced, 0 at the end, -1 on error.  Pointers stay
 * valid until the next call.  frag_* describe the chunk-relative fragment table (entry 0 starts at 0). */
int  blamm_fasta_next(blamm_fasta* f, uint64_t payload, uint64_t halo, const char** chars, uint64_t* n_total,
                      uint64_t* n_payload, uint64_t* stream_start, const uint64_t** frag_start,
                      const uint64_t** frag_seq, const uint64_t** frag_pos, uint64_t* n_frag);
int  blamm_fasta_num_sequences(const blamm_fasta* f);
const char* blamm_fasta_sequence_name(const blamm_fasta* f, int idx);
int  blamm_fasta_counts(const blamm_fasta* f, uint64_t counts[4]);
/* 2-bit packer for b200scan_submit_packed -- replaces the FP32 one-hot fill of SeqMatrix::getNextSeqMatrix
 * (sequence.cpp:306-337: 16 bytes per character) by 0.375 byte per character, in the layout b200scan.h documents:
 * character i -> bits 2(i % 16) of codes2[i / 16] (A0 C1 G2 T3, either case) and bit i % 32 of zero_mask[i / 32]
 * (set = contributes 0: lower case unless fold_lower -- the reference's fill only recognises upper case,
 * sequence.cpp:312-319 -- and any byte outside ACGTacgt; padding behind n: code 0, bit set).  codes2 holds
 * ceil(n / 16) words, zero_mask ceil(n / 32).  blamm_pack_ascii packs any buffer on the calling thread,
 * blamm_fasta_pack the chunk the last blamm_fasta_next returned, on the stream's parser threads.
 * Return 1 if one of the n characters has its zero bit set, 0 if none, -1 on error. */
int  blamm_pack_ascii(const char* chars, uint64_t n, int fold_lower, uint32_t* codes2, uint32_t* zero_mask);
int  blamm_fasta_pack(blamm_fasta* f, int fold_lower, uint32_t* codes2, uint32_t* zero_mask);

/* Score column of the occurrence file: the text `ostream << float` produces in the reference's writer (pwmscan.cpp:94),
 * i.e. "%g" with 6 significant digits.  dst needs 32 bytes; returns the length (no terminator is written). */
int  blamm_format_score(float score, char* dst);

#ifdef __cplusplus
}
#endif
#endif
