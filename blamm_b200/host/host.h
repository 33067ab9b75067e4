// Host-side model of the `blamm scan` path: inputs (settings.cnf, .dict, JASPAR, histogram files), the
// background-corrected PWM matrix and thresholds, the filtered sequence stream and the occurrence writer.
// Behaviour (file formats, formulas, messages) follows biointec/blamm so the CLI is a drop-in; the structure
// is new: one flat FP32 matrix + per-column metadata for the device, a chunked stream with a global
// fragment table instead of 250,000-character SeqBlocks, and no BLAS anywhere.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace blamm {

// settings.cnf, read from the current directory (reference: settings.cpp:34-74)
struct Settings {
    size_t matrix_S_w = 250, matrix_S_h = 1000, matrix_P_tile_min_zero_area = 64 * 64, flushOutput = 100000;
    float pseudocount = 0.25f;
    bool defaultVal = true;
    Settings();                         // reads ./settings.cnf if present
    explicit Settings(const std::string& path);
    void print() const;                 // reference: settings.cpp:76-88
private:
    void load(const std::string& path);
};

// One motif = one column of P once reverse complements have been added (reference: class Motif, motif.h:152-334)
struct Motif {
    std::string name;
    std::vector<std::array<uint64_t, 4>> pfm;      // counts per position, ACGT
    std::vector<std::array<float, 4>> pwm;         // log2-odds, filled by computePWM
    float threshold = 0.0f;
    bool revComp = false;

    size_t size() const { return pfm.size(); }
    std::string baseName() const;                  // name without a trailing __NN (motif.h:256-267)
    bool isPermutation() const;                    // motif.h:273-281
    void computePWM(const std::array<uint64_t, 4>& bgCounts, float pseudo);   // motif.cpp:194-223
    void reverseComplement();                      // motif.cpp:267-286
    float maxScore() const;                        // motif.cpp:241-252
    float minScore() const;                        // motif.cpp:254-265
};

// Score histogram file (reference: class ScoreHistogram, motif.h:36-146; motif.cpp:47-132)
struct ScoreHistogram {
    std::vector<uint64_t> counts;
    float minScore = 0, maxScore = 0, width = 0;
    ScoreHistogram() = default;
    ScoreHistogram(float mn, float mx, size_t bins);
    void setNumObservations(float score, uint64_t count);
    void addObservation(float score);
    float scoreCutoff(float pvalue) const;
    void load(const std::string& dir, const std::string& base);
    void writeGNUPlot(const std::string& dir, const std::string& base, const std::string& label) const;
};

class MotifSet {
public:
    std::vector<Motif> motifs;
    void load(const std::string& filename, bool loadPermutations);     // motif.cpp:323-338 (JASPAR or cluster-buster)
    void addReverseComplements();                                       // motif.cpp:439-449
    size_t maxLen() const;
    // PWMs for a background + the flat matrix P (column major, ld = 4*maxLen, zero padded; motif.cpp:542-564)
    void generateMatrix(const std::array<uint64_t, 4>& bgCounts, float pseudo);
    const std::vector<float>& P() const { return P_; }
    int ldp() const { return (int)(4 * maxLen()); }
    std::vector<int32_t> colLen() const;
    std::vector<float> colThr() const;
    // theoretical score spectrum of one motif (motif.cpp:151-192) into a histogram (hist.cpp:162-175)
    static void theoreticalHistogram(const Motif& m, const std::array<float, 4>& bgProb, size_t numBins,
                                     uint64_t maxLength, ScoreHistogram& out);
private:
    std::vector<float> P_;
};

// One manifest group (reference: class Species, species.h:33-217)
struct Species {
    std::string name;
    std::vector<std::string> files;                // lexicographic order (std::set in the reference, species.h:36)
    std::array<uint64_t, 4> nuclCounts{{0, 0, 0, 0}};
    uint64_t totSeqLen = 0;
    std::vector<std::string> seqNames;
    std::array<float, 4> nuclProb(float pseudo) const;     // species.cpp:63-71
    void printNuclProb(float pseudo) const;                // species.cpp:73-82
};

struct SpeciesSet {
    std::vector<Species> species;
    void loadDict(const std::string& filename);            // species.cpp:190-207
    void writeDict(const std::string& filename) const;     // species.cpp:176-188
    void addFile(const std::string& group, const std::string& fasta);   // species.cpp:149-164
};

// A fragment of the filtered stream: a maximal run of ACGTacgt characters that is contiguous in one record
// (reference: the SeqBlock::block2seq markers, sequence.cpp:35-50, 81-88).
struct Fragment { uint64_t streamPos; uint64_t seqIdx; uint64_t seqPos; };

// Filtered stream of one group, produced chunk by chunk (reference: FastaBatch, sequence.cpp:123-293).
// The files are mmap'ed and parsed in waves of byte segments, one segment per thread; a segment is parsed without
// knowing the record it starts in (record index and in-record position relative to a carry-in), and a short serial
// stitch resolves the carries, merges fragments across segment borders and applies the maxFiltered cut.
class FastaStream {
public:
    FastaStream(const std::vector<std::string>& files, uint64_t maxFiltered = UINT64_MAX);
    ~FastaStream();
    FastaStream(const FastaStream&) = delete;
    FastaStream& operator=(const FastaStream&) = delete;
    // Parser threads (>= 1) and, for tests, a forced segment size in bytes (0 = sized from the request).
    void setParallel(unsigned threads, size_t segmentBytes = 0);
    // Next chunk: up to `payload` new characters followed by up to `halo` characters that will open the next
    // chunk.  Returns false when the stream is exhausted.  chunk.chars stays valid until the next call.
    struct Chunk {
        const char* chars = nullptr;
        uint64_t nTotal = 0, nPayload = 0;
        uint64_t streamStart = 0;                  // global stream position of chars[0]
        std::vector<uint64_t> fragStarts;          // chunk-relative starts of fragments beginning inside (0, nTotal)
        std::vector<Fragment> frags;               // chunk-relative fragment table, frags[0].streamPos == 0
    };
    bool next(uint64_t payload, uint64_t halo, Chunk& out);
    void copyChunk(const Chunk& c, char* dst);     // c.chars[0, nTotal) -> dst on the parser threads
    // c.chars[0, nTotal) -> 2-bit codes + zero mask (see packAscii) on the parser threads; true if a character contributes 0
    bool packChunk(const Chunk& c, bool foldLower, uint32_t* codes2, uint32_t* zmask);
    const std::vector<Fragment>& fragments() const { return frags_; }
    const std::vector<std::string>& seqNames() const { return names_; }
    uint64_t filteredLength() const { return streamLen_; }
    // stream position -> (record index, position in record)   (SeqBlock::getSeqPos, sequence.cpp:52-66)
    void locate(uint64_t streamPos, uint64_t& seqIdx, uint64_t& seqPos) const;
    // nucleotide counts of everything consumed so far (dict module, species.cpp:32-61)
    const std::array<uint64_t, 4>& counts() const { return counts_; }
    struct Segment;                                // one parsed byte range (sequence.cpp)
    struct Pool;                                   // the parser threads (sequence.cpp)
private:
    bool fill(uint64_t want);                      // append filtered characters until `want` are buffered
    bool openNext();
    void wave(uint64_t need);                      // parse the next byte range of the current file
    void stitch(Segment& s, const char* p, size_t n, bool midLine);
    void reserveBuf(size_t n);
    std::vector<std::string> files_;
    size_t fileIdx_ = 0;
    const char* map_ = nullptr; size_t mapLen_ = 0, mapPos_ = 0; int fd_ = -1;
    bool posMidLine_ = false;                      // mapPos_ lies inside a sequence line (a long line was split)
    uint64_t maxFiltered_, streamLen_ = 0;
    uint64_t curSeqLen_ = 0;                       // position inside the current record
    bool haveLast_ = false; uint64_t lastSeq_ = 0, lastPosPlus1_ = 0;
    uint64_t pendingDrop_ = 0;
    // filtered characters [bufStart_, bufStart_ + bufLen_) live at buf_[bufHead_ ...]; plain malloc'ed storage so
    // that growing it does not zero-fill what the copy threads overwrite anyway
    char* buf_ = nullptr; size_t bufCap_ = 0, bufHead_ = 0, bufLen_ = 0; uint64_t bufStart_ = 0;
    std::vector<Fragment> frags_;
    std::vector<std::string> names_;
    std::array<uint64_t, 4> counts_{{0, 0, 0, 0}};
    bool eof_ = false;
    unsigned threads_ = 1; size_t forcedSegment_ = 0;
    std::vector<std::unique_ptr<Segment>> segs_;   // reused from wave to wave
    std::unique_ptr<Pool> pool_;
};

// Host twin of the device packer (csrc/pack.cuh) for b200scan_submit_packed: character i -> bits 2(i % 16) of codes2[i / 16]
// (A0 C1 G2 T3, either case) and bit i % 32 of zmask[i / 32] ("contributes zero": lower case unless foldLower -- the
// reference's one-hot fill only recognises upper case, sequence.cpp:312-319 -- and any byte outside ACGTacgt; the padding
// behind n in the last words is code 0 / zero 1).  codes2 holds ceil(n / 16) words, zmask ceil(n / 32).  Returns true if
// one of the n characters has its zero bit set.  Replaces the 16 bytes of FP32 one-hot per character that
// SeqMatrix::getNextSeqMatrix writes (sequence.cpp:306-337) by 0.375 byte.
bool packAscii(const char* chars, uint64_t n, bool foldLower, uint32_t* codes2, uint32_t* zmask);

// "%g" with 6 significant digits == ostream << float (pwmscan.cpp:94)
int formatScore(char* dst, float v);

// CLI modules (reference: blstools.cpp:63-101)
int runDict(int argc, char** argv);
int runHist(int argc, char** argv);
int runScan(int argc, char** argv);

} // namespace blamm
