/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's `blamm scan` hot path
 * (biointec/blamm), written from the reference's behaviour, each function citing the file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (libb200scan.so, blamm-b200) never does.
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this oracle is
 * pinned against the reference ITSELF, compiled from /root/reference into oracle/_ref/ (build_ref.sh):
 * tests/test_oracle_golden.py compares it hit-for-hit and bit-for-bit (scores, thresholds, P) with
 * oracle/_ref/refdump on the example and on seeded synthetic inputs, and tests/golden/ holds fixtures
 * produced by oracle/_ref/blamm (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -shared -fPIC oracle.c -o liboracle.so -lm     (oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------
 * Motif::PFM2PWM (motif.cpp:194-223).  pfm: L rows of 4 counts (A,C,G,T), already reverse-complemented for a
 * reverse-complement column (Motif::revCompl, motif.cpp:267-286); bg: the group's nucleotide counts;
 * revcomp != 0 complements the background (motif.cpp:204-207).  All arithmetic in float like the reference.
 * ---------------------------------------------------------------------------------------------------- */
void oracle_pfm2pwm(const uint64_t* pfm, int L, const uint64_t bg[4], float pseudo, int revcomp, float* pwm)
{
    float bgTot = (float)(bg[0] + bg[1] + bg[2] + bg[3]);
    bgTot += 4.0f * pseudo;
    float bgProb[4];
    for (int i = 0; i < 4; i++) bgProb[i] = ((float)bg[i] + pseudo) / bgTot;
    if (revcomp) {
        float t = bgProb[0]; bgProb[0] = bgProb[3]; bgProb[3] = t;
        t = bgProb[1]; bgProb[1] = bgProb[2]; bgProb[2] = t;
    }
    for (int i = 0; i < L; i++) {
        float tot = (float)(pfm[4 * i] + pfm[4 * i + 1] + pfm[4 * i + 2] + pfm[4 * i + 3]);
        tot += 4.0f * pseudo;
        for (int j = 0; j < 4; j++) {
            float ppm = ((float)pfm[4 * i + j] + pseudo) / tot;
            pwm[4 * i + j] = log2f(ppm / bgProb[j]);
        }
    }
}

/* Species::getNuclProbabilities (species.cpp:63-71) */
void oracle_bg_prob(const uint64_t bg[4], float pseudo, float out[4])
{
    float bgTot = (float)(bg[0] + bg[1] + bg[2] + bg[3]);
    bgTot += 4.0f * pseudo;
    for (int i = 0; i < 4; i++) out[i] = ((float)bg[i] + pseudo) / bgTot;
}

/* Motif::getMaxScore / getMinScore (motif.cpp:241-265): per position max/min, summed in float in order. */
float oracle_max_score(const float* pwm, int L)
{
    float s = 0.0f;
    for (int i = 0; i < L; i++) {
        float a = fmaxf(pwm[4 * i], pwm[4 * i + 1]), b = fmaxf(pwm[4 * i + 2], pwm[4 * i + 3]);
        s += fmaxf(a, b);
    }
    return s;
}
float oracle_min_score(const float* pwm, int L)
{
    float s = 0.0f;
    for (int i = 0; i < L; i++) {
        float a = fminf(pwm[4 * i], pwm[4 * i + 1]), b = fminf(pwm[4 * i + 2], pwm[4 * i + 3]);
        s += fminf(a, b);
    }
    return s;
}

/* ------------------------------------------------------------------------------------------------------
 * Theoretical score histogram = Motif::computeTheoreticalSpectrum (motif.cpp:151-192) followed by
 * Histogram::generateTheoreticalHist -> ScoreHistogram::setNumObservations (hist.cpp:162-175,
 * motif.h:109-115).  The reference keeps std::map<int,float> per position and walks it in ascending key
 * order; a dense array walked in ascending index order adds the same floats in the same order.
 * counts (numBins entries) must be zero-initialised by the caller.  Returns 0, or -1 on allocation failure.
 * ---------------------------------------------------------------------------------------------------- */
int oracle_theoretical_hist(const float* pwm, int L, const float bgprob[4], int numBins, uint64_t maxLength,
                            uint64_t* counts)
{
    float minS = oracle_min_score(pwm, L), maxS = oracle_max_score(pwm, L);
    float range = maxS - minS;
    float a = (float)numBins / range;
    float b = -a * minS / (float)L;
    int* w = (int*)malloc(sizeof(int) * 4 * L);
    if (!w) return -1;
    long lo = 0, hi = 0;   /* reachable integer score range */
    for (int p = 0; p < L; p++) {
        int mn = 0, mx = 0;
        for (int k = 0; k < 4; k++) {
            w[4 * p + k] = (int)round(a * pwm[4 * p + k] + b);     /* float expr promoted to double by round() */
            if (k == 0 || w[4 * p + k] < mn) mn = w[4 * p + k];
            if (k == 0 || w[4 * p + k] > mx) mx = w[4 * p + k];
        }
        lo += mn; hi += mx;
    }
    /* offsets so that every partial sum is indexable */
    long plo = 0, phi = 0, gmin = 0, gmax = 0;
    for (int p = 0; p < L; p++) {
        int mn = w[4 * p], mx = w[4 * p];
        for (int k = 1; k < 4; k++) { if (w[4 * p + k] < mn) mn = w[4 * p + k]; if (w[4 * p + k] > mx) mx = w[4 * p + k]; }
        plo += mn; phi += mx;
        if (plo < gmin) gmin = plo;
        if (phi > gmax) gmax = phi;
    }
    (void)lo; (void)hi;
    long span = gmax - gmin + 1;
    float* cur = (float*)calloc(span, sizeof(float));
    float* nxt = (float*)calloc(span, sizeof(float));
    char* curSet = (char*)calloc(span, 1);
    char* nxtSet = (char*)calloc(span, 1);
    if (!cur || !nxt || !curSet || !nxtSet) { free(w); free(cur); free(nxt); free(curSet); free(nxtSet); return -1; }
    for (int k = 0; k < 4; k++) { long s = w[k] - gmin; cur[s] += bgprob[k]; curSet[s] = 1; }        /* motif.cpp:173-174 */
    for (int p = 1; p < L; p++) {
        memset(nxt, 0, span * sizeof(float)); memset(nxtSet, 0, span);
        for (long s = 0; s < span; s++) {                                                             /* motif.cpp:177-184 */
            if (!curSet[s]) continue;
            for (int k = 0; k < 4; k++) {
                long t = s + w[4 * p + k];
                nxt[t] += cur[s] * bgprob[k];
                nxtSet[t] = 1;
            }
        }
        float* tf = cur; cur = nxt; nxt = tf;
        char* tc = curSet; curSet = nxtSet; nxtSet = tc;
    }
    /* spectrum in the original score range (motif.cpp:187-191), then setNumObservations in ascending score order */
    float width = (maxS - minS) / (float)numBins;                                                     /* motif.h:66 */
    for (long s = 0; s < span; s++) {
        if (!curSet[s]) continue;
        float score = ((float)(s + gmin) - L * b) / a;
        uint64_t cnt = (uint64_t)((float)maxLength * cur[s]);        /* size_t * float -> float, truncated to size_t (hist.cpp:172) */
        int bin = (int)((score - minS) / width);
        if (bin < 0) bin = 0;
        if (bin > numBins - 1) bin = numBins - 1;
        counts[bin] = cnt;                                           /* store, not accumulate (motif.h:114) */
    }
    free(w); free(cur); free(nxt); free(curSet); free(nxtSet);
    return 0;
}

/* ScoreHistogram::getScoreCutoff (motif.cpp:47-69); width as in loadHistogram (motif.cpp:119). */
float oracle_score_cutoff(const uint64_t* counts, int numBins, float minScore, float maxScore, float pvalue)
{
    float width = (maxScore - minScore) / (float)numBins;
    double totObs = 0.0;
    for (int i = 0; i < numBins; i++) totObs += (double)counts[i];
    double bestObs = pvalue * totObs;
    double curr = bestObs;
    for (long i = numBins - 1; i >= 0; i--) {
        if ((double)counts[i] < curr) curr -= (double)counts[i];
        else {
            double frac = curr / (double)counts[i];
            float cutoffi = (float)(frac * i + (1.0 - frac) * (i + 1));
            return cutoffi * width + minScore;
        }
    }
    return maxScore;
}

/* ------------------------------------------------------------------------------------------------------
 * FASTA -> filtered stream.  Restates FastaBatch::getNextLine / filterLine / appendNextBlock and
 * SeqBlock::append / isContiguous (sequence.cpp:35-50, 81-88, 142-250) for one file image appended to a
 * running state: characters ACGTacgt are kept, anything else ends a fragment but advances the in-record
 * position; a new fragment starts whenever the kept character is not contiguous (same record, consecutive
 * position) with the previous kept character.
 * state[0] = number of records so far, state[1] = position inside the current record, state[2] = stream length
 * so far, state[3] = last kept (record) , state[4] = last kept position + 1 (0 = none), state[5] = n fragments.
 * Outputs are appended: stream (chars), frag_start/frag_seq/frag_pos (one entry per fragment).
 * name_off receives the byte offset (in `file`) of each '>' line (names are cut by the caller).
 * Returns 0, -1 if a sequence line precedes the first header (sequence.cpp:167-168), -2 on capacity overflow.
 * ---------------------------------------------------------------------------------------------------- */
int oracle_fasta_filter(const char* file, uint64_t n, uint64_t* state, char* stream, uint64_t stream_cap,
                        uint64_t* frag_start, uint64_t* frag_seq, uint64_t* frag_pos, uint64_t frag_cap,
                        uint64_t* name_off, uint64_t name_cap, uint64_t* n_names, uint64_t max_filtered)
{
    uint64_t i = 0;
    while (i < n) {
        uint64_t e = i;
        while (e < n && file[e] != '\n') e++;            /* std::getline */
        uint64_t len = e - i;
        if (len == 0) { i = e + 1; continue; }           /* skip empty lines (sequence.cpp:150-151) */
        if (file[i] == '>') {                            /* header (sequence.cpp:154-162) */
            if (*n_names >= name_cap) return -2;
            name_off[(*n_names)++] = i;
            state[0]++; state[1] = 0;
            i = e + 1; continue;
        }
        if (state[0] == 0) return -1;
        for (uint64_t k = i; k < e; k++) {
            char c = file[k];
            int valid = (c == 'A' || c == 'a' || c == 'C' || c == 'c' || c == 'G' || c == 'g' || c == 'T' || c == 't');
            if (valid && state[2] < max_filtered) {      /* _maxFiltSeqLen cap (sequence.cpp:219-220) */
                int contiguous = (state[4] != 0 && state[3] == state[0] - 1 && state[4] == state[1]);
                if (!contiguous) {
                    if (state[5] >= frag_cap) return -2;
                    frag_start[state[5]] = state[2]; frag_seq[state[5]] = state[0] - 1; frag_pos[state[5]] = state[1];
                    state[5]++;
                }
                if (state[2] >= stream_cap) return -2;
                stream[state[2]++] = c;
                state[3] = state[0] - 1; state[4] = state[1] + 1;
            }
            state[1]++;                                  /* every character advances the record position (sequence.cpp:208) */
        }
        i = e + 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * The scan of one filtered stream: PWMScan::scanThreadBLAS + extractOccurrences (pwmscan.cpp:223-278,
 * 104-133) with SeqMatrix::getNextSeqMatrix's one-hot (sequence.cpp:306-337) and BLAS sgemm on it restated
 * as what they compute: for every stream position g and column c
 *     score = ((0 + w_0) + w_1) + ... + w_{L-1},   w_j = P[c*ldp + 4j + code(stream[g+j])]  in FP32,
 * where only UPPER-case letters have a one-hot entry (lower case contributes 0, lower_fold == 0) -- the
 * in-order chain is what a BLAS micro-kernel does on a one-hot operand (verified bit-exact against
 * OpenBLAS through oracle/_ref/refdump); a hit is kept iff !(score < thr) and the window lies inside one
 * fragment (SeqBlock::getRemainingSeqLen, sequence.cpp:68-79).  The reference's block size h*w and its
 * maxLen-1 overlap only decide WHICH block scores a window, never the value, so one pass over the whole
 * stream is equivalent.  lower_fold != 0 gives the naive path's case-insensitive scoring (motif.cpp:138-149).
 * frag_start: ascending stream positions where a fragment begins (first is 0).
 * Hits (stream position, column, score) are written up to cap; returns the total number found.
 * ---------------------------------------------------------------------------------------------------- */
static int code_of(char c, int lower_fold)
{
    switch (c) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        case 'a': return lower_fold ? 0 : 4; case 'c': return lower_fold ? 1 : 4;
        case 'g': return lower_fold ? 2 : 4; case 't': return lower_fold ? 3 : 4;
        default: return 4;
    }
}

uint64_t oracle_scan_stream(const char* stream, uint64_t n, uint64_t n_payload, const uint64_t* frag_start,
                            uint64_t n_frag, const float* P, int ldp, int n_cols, const int32_t* col_len,
                            const float* thr, int lower_fold, uint64_t* hit_pos, uint32_t* hit_col,
                            float* hit_score, uint64_t cap)
{
    uint64_t nh = 0, f = 0;
    uint8_t* code = (uint8_t*)malloc(n + 1);
    if (!code) return (uint64_t)-1;
    for (uint64_t g = 0; g < n; g++) code[g] = (uint8_t)code_of(stream[g], lower_fold);
    for (uint64_t g = 0; g < n_payload && g < n; g++) {
        while (f + 1 < n_frag && frag_start[f + 1] <= g) f++;
        uint64_t frag_end = (f + 1 < n_frag) ? frag_start[f + 1] : n;
        uint64_t remaining = frag_end - g;
        uint64_t avail = n - g;
        for (int c = 0; c < n_cols; c++) {
            int L = col_len[c];
            const float* w = P + (size_t)c * ldp;
            float s = 0.0f;
            int lim = (uint64_t)L < avail ? L : (int)avail;      /* beyond the block end S is all zero */
            for (int j = 0; j < lim; j++) {
                int k = code[g + j];
                if (k < 4) s += w[4 * j + k];
            }
            if (s < thr[c]) continue;                            /* pwmscan.cpp:116 */
            if ((uint64_t)L > remaining) continue;               /* pwmscan.cpp:124 */
            if (nh < cap) { hit_pos[nh] = g; hit_col[nh] = (uint32_t)c; hit_score[nh] = s; }
            nh++;
        }
    }
    free(code);
    return nh;
}

/* ------------------------------------------------------------------------------------------------------
 * Empirical score histogram = Histogram::histThread + extractObsScore + ScoreHistogram::addObservation
 * (hist.cpp:70-140, motif.h:62-66, 96-102): every window of the (already truncated) filtered stream that lies
 * inside one fragment adds one observation to bin clamp(int((score - min) / width), 0, bins-1) of its column,
 * width = (max - min) / (float)bins.  Scores as in oracle_scan_stream.  counts: n_cols * num_bins, zeroed by the caller.
 * ---------------------------------------------------------------------------------------------------- */
int oracle_empirical_hist(const char* stream, uint64_t n, const uint64_t* frag_start, uint64_t n_frag, const float* P,
                          int ldp, int n_cols, const int32_t* col_len, const float* mins, const float* maxs,
                          int num_bins, int lower_fold, uint64_t* counts)
{
    uint64_t f = 0;
    uint8_t* code = (uint8_t*)malloc(n + 1);
    if (!code) return -1;
    for (uint64_t g = 0; g < n; g++) code[g] = (uint8_t)code_of(stream[g], lower_fold);
    for (uint64_t g = 0; g < n; g++) {
        while (f + 1 < n_frag && frag_start[f + 1] <= g) f++;
        uint64_t frag_end = (f + 1 < n_frag) ? frag_start[f + 1] : n;
        uint64_t remaining = frag_end - g;
        for (int c = 0; c < n_cols; c++) {
            int L = col_len[c];
            if ((uint64_t)L > remaining) continue;                   /* hist.cpp:87-88 */
            const float* w = P + (size_t)c * ldp;
            float s = 0.0f;
            for (int j = 0; j < L; j++) { int k = code[g + j]; if (k < 4) s += w[4 * j + k]; }
            float width = (maxs[c] - mins[c]) / (float)num_bins;
            int bin = (int)((s - mins[c]) / width);
            if (bin < 0) bin = 0;
            if (bin > num_bins - 1) bin = num_bins - 1;
            counts[(size_t)c * num_bins + bin]++;
        }
    }
    free(code);
    return 0;
}

/* Same scoring for explicit (position, column) pairs -- used by tests to check individual scores. */
void oracle_score_at(const char* stream, uint64_t n, const float* P, int ldp, const int32_t* col_len, int lower_fold,
                     const uint64_t* pos, const uint32_t* col, uint64_t m, float* out)
{
    for (uint64_t i = 0; i < m; i++) {
        int L = col_len[col[i]];
        const float* w = P + (size_t)col[i] * ldp;
        float s = 0.0f;
        for (int j = 0; j < L && pos[i] + j < n; j++) {
            int k = code_of(stream[pos[i] + j], lower_fold);
            if (k < 4) s += w[4 * j + k];
        }
        out[i] = s;
    }
}
