// Species dictionary and the filtered sequence stream.
#include "host.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <fstream>
#include <iostream>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <thread>
#include <cstdlib>

#include <fcntl.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>      // the AVX2 / BMI2 packer is compiled with a target attribute and chosen at run time
#endif
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace blamm {

// ---------------------------------------------------------------------------------------------------------
// Species / dictionary  (file format: reference species.cpp:84-143, 176-207)
// ---------------------------------------------------------------------------------------------------------
std::array<float, 4> Species::nuclProb(float pseudo) const
{
    float tot = (float)(nuclCounts[0] + nuclCounts[1] + nuclCounts[2] + nuclCounts[3]);
    tot += 4.0f * pseudo;
    std::array<float, 4> p;
    for (int i = 0; i < 4; i++) p[i] = ((float)nuclCounts[i] + pseudo) / tot;
    return p;
}

void Species::printNuclProb(float pseudo) const
{
    const auto p = nuclProb(pseudo);
    const auto old = std::cout.precision();
    std::cout.precision(3);
    std::cout << " [A: " << 100.0f * p[0] << "%, C: " << 100.0f * p[1] << "%, G: " << 100.0f * p[2]
              << "%, T: " << 100.0f * p[3] << "%]\n";
    std::cout.precision(old);
}

static void expectKey(std::istream& in, const char* key, const std::string& file)
{
    std::string k;
    in >> k;
    if (k != key) throw std::runtime_error("Malformed dictionary " + file + ": expected " + key + ", found '" + k + "'");
}

void SpeciesSet::loadDict(const std::string& filename)
{
    std::ifstream in(filename);
    if (!in) throw std::runtime_error("Cannot open file: " + filename);
    size_t n = 0;
    expectKey(in, "NUM_SPECIES", filename); in >> n;
    for (size_t i = 0; i < n; i++) {
        Species s;
        size_t nFiles = 0, nSeq = 0;
        expectKey(in, "SPECIES", filename); in >> s.name;
        expectKey(in, "NUM_FASTA_FILES", filename); in >> nFiles;
        std::set<std::string> files;
        for (size_t f = 0; f < nFiles; f++) { std::string t; in >> t; files.insert(t); }
        s.files.assign(files.begin(), files.end());
        expectKey(in, "TOT_SEQ_LENGTH", filename); in >> s.totSeqLen;
        expectKey(in, "NUCL_COUNT_ACGT", filename);
        in >> s.nuclCounts[0] >> s.nuclCounts[1] >> s.nuclCounts[2] >> s.nuclCounts[3];
        expectKey(in, "NUM_SEQUENCES", filename); in >> nSeq;
        s.seqNames.resize(nSeq);
        for (auto& nm : s.seqNames) in >> nm;
        species.push_back(std::move(s));
    }
}

void SpeciesSet::writeDict(const std::string& filename) const
{
    std::ofstream out(filename);
    if (!out) throw std::runtime_error("Cannot write to file: " + filename);
    out << "NUM_SPECIES\t" << species.size() << "\n";
    for (const auto& s : species) {
        out << "SPECIES\t" << s.name << "\n" << "NUM_FASTA_FILES\t" << s.files.size() << "\n";
        for (const auto& f : s.files) out << f << "\n";
        out << "TOT_SEQ_LENGTH\t" << s.totSeqLen << "\n";
        out << "NUCL_COUNT_ACGT\t" << s.nuclCounts[0] << "\t" << s.nuclCounts[1] << "\t" << s.nuclCounts[2] << "\t"
            << s.nuclCounts[3] << "\n";
        out << "NUM_SEQUENCES\t" << s.seqNames.size() << "\n";
        for (const auto& nm : s.seqNames) out << nm << "\n";
    }
}

void SpeciesSet::addFile(const std::string& group, const std::string& fasta)
{
    auto it = std::find_if(species.begin(), species.end(), [&](const Species& s) { return s.name == group; });
    if (it == species.end()) { species.push_back(Species()); it = species.end() - 1; it->name = group; }
    // keep the file list a sorted set (reference: std::set<std::string>, species.h:36)
    auto pos = std::lower_bound(it->files.begin(), it->files.end(), fasta);
    if (pos == it->files.end() || *pos != fasta) it->files.insert(pos, fasta);
}

// ---------------------------------------------------------------------------------------------------------
// FastaStream
//
// Semantics (reference sequence.cpp:142-250): lines are split at '\n'; empty lines are skipped; a line starting
// with '>' opens a record whose name is the first whitespace-delimited token after '>'; every other line is
// sequence: characters ACGTacgt are kept, any other character ends the current fragment, and EVERY character
// advances the position inside the record.  Kept characters of all records and files are concatenated; a new
// fragment starts whenever a kept character does not directly follow the previous kept one in the same
// record.  The stream stops after maxFiltered kept characters (the dictionary's TOT_SEQ_LENGTH).
//
// Structure: the reference parses line by line under a mutex (sequence.cpp:274-293).  Here a *wave* cuts the next
// byte range of the mmap'ed file into one segment per thread (at line starts; a line longer than 4 KB is split in
// the middle), every segment is parsed independently into its own buffer with record indices and in-record
// positions relative to an unknown carry-in, and a serial stitch over the segments' few fragments and names
// resolves the carries.  The kept characters are then copied into the stream buffer in parallel.
// ---------------------------------------------------------------------------------------------------------
static const struct ValidTable {
    uint8_t v[256];
    ValidTable() { std::memset(v, 4, sizeof v);
                   v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; }
} kValid;

// First index >= i (and <= len) whose character is not KEPT (ACGTacgt) / not dropped: 16 characters per step with
// SSE2 (baseline x86-64), bytes elsewhere.
#if defined(__SSE2__)
static inline unsigned keptMask16(const char* p)
{
    const __m128i u = _mm_or_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p)), _mm_set1_epi8(0x20));
    const __m128i k = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('a')), _mm_cmpeq_epi8(u, _mm_set1_epi8('c'))),
                                   _mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('g')), _mm_cmpeq_epi8(u, _mm_set1_epi8('t'))));
    return (unsigned)_mm_movemask_epi8(k);
}
#endif
template <bool KEPT>
static inline size_t runEnd(const char* line, size_t i, size_t len)
{
#if defined(__SSE2__)
    while (i + 16 <= len) {
        const unsigned m = KEPT ? ~keptMask16(line + i) & 0xFFFFu : keptMask16(line + i);    // bits that end the run
        if (m) return i + (size_t)__builtin_ctz(m);
        i += 16;
    }
#endif
    while (i < len && (kValid.v[(uint8_t)line[i]] != 4) == KEPT) i++;
    return i;
}

// A, C, G, T counts of a run of kept characters.  (c >> 1) & 3 is 0 for Aa, 1 for Cc, 3 for Gg, 2 for Tt, so two bit
// planes of eight characters at a time give the C, G and T counts in byte lanes; A is the remainder.
static void countRun(const char* s, size_t n, uint64_t cnt[4])
{
    const uint64_t kOnes = 0x0101010101010101ull;
    uint64_t c = 0, g = 0, t = 0;
    size_t i = 0;
    while (i + 8 <= n) {
        uint64_t ac = 0, ag = 0, at = 0;
        const size_t stop = std::min(n - (n - i) % 8, i + 8 * 255);
        for (; i < stop; i += 8) {
            uint64_t w;
            std::memcpy(&w, s + i, 8);
            const uint64_t b1 = (w >> 1) & kOnes, b2 = (w >> 2) & kOnes;
            ac += b1 & ~b2; ag += b1 & b2; at += b2 & ~b1;
        }
        // horizontal sums of the byte lanes (each lane <= 255)
        auto hsum = [](uint64_t v) {
            v = (v & 0x00FF00FF00FF00FFull) + ((v >> 8) & 0x00FF00FF00FF00FFull);
            v = (v & 0x0000FFFF0000FFFFull) + ((v >> 16) & 0x0000FFFF0000FFFFull);
            return (v & 0xFFFFFFFFull) + (v >> 32);
        };
        c += hsum(ac); g += hsum(ag); t += hsum(at);
    }
    for (; i < n; i++) {
        const unsigned k = ((unsigned char)s[i] >> 1) & 3;
        c += k == 1; g += k == 3; t += k == 2;
    }
    cnt[0] += n - c - g - t; cnt[1] += c; cnt[2] += g; cnt[3] += t;
}

// One parsed byte range.  Records are numbered relative to the segment: 0 = the record that was open when the segment
// began (its index and the position reached in it are the carry-in), k = the k-th header line inside the segment.
struct FastaStream::Segment {
    struct Frag { uint64_t off; uint32_t rec; uint64_t pos; };     // off = offset into `chars`
    std::unique_ptr<char[]> chars; size_t cap = 0, nKept = 0;
    std::vector<Frag> frags;
    std::vector<std::string> names;
    uint64_t seqLenEnd = 0;              // in-record position at the end (to be added to the carry-in if names is empty)
    bool haveLast = false; uint32_t lastRec = 0; uint64_t lastPosPlus1 = 0;
    uint64_t counts[4] = {0, 0, 0, 0};
    bool seqBeforeHeader = false;        // a non-empty sequence line belongs to record 0
    const char* src = nullptr; size_t srcLen = 0; bool midLine = false;
    size_t dstOff = 0;                   // where the kept characters go in the stream buffer (set by the stitch)

    // limit = how many characters may still be kept; like the reference, the cut is tested before every line
    void parse(const char* p, size_t n, bool startsMidLine, uint64_t limit)
    {
        if (cap < n) { chars.reset(new char[n]); cap = n; }
        nKept = 0; frags.clear(); names.clear(); seqLenEnd = 0; haveLast = false; lastRec = 0; lastPosPlus1 = 0;
        counts[0] = counts[1] = counts[2] = counts[3] = 0; seqBeforeHeader = false;
        char* dst = chars.get();
        uint32_t rec = 0;
        uint64_t seqLen = 0;
        size_t at = 0;
        bool mid = startsMidLine;
        while (at < n && nKept < limit) {
            const char* line = p + at;
            const char* nl = static_cast<const char*>(memchr(line, '\n', n - at));
            const size_t len = nl ? (size_t)(nl - line) : n - at;
            at += len + (nl ? 1 : 0);
            const bool lineStart = !mid;
            mid = false;
            if (len == 0) continue;
            if (lineStart && line[0] == '>') {
                seqLen = 0; rec++;
                size_t b = 1;
                while (b < len && isspace((unsigned char)line[b])) b++;
                size_t e = b;
                while (e < len && !isspace((unsigned char)line[e])) e++;
                names.emplace_back(line + b, e - b);
                continue;
            }
            if (rec == 0) seqBeforeHeader = true;
            size_t i = 0;
            while (i < len) {
                i = runEnd<false>(line, i, len);                                 // a run of dropped characters
                const size_t j = runEnd<true>(line, i, len);
                if (j == i) break;
                size_t run = j - i;
                if (nKept + run > limit) run = (size_t)(limit - nKept);
                if (run) {
                    const uint64_t pos = seqLen + i;
                    if (!(haveLast && lastRec == rec && lastPosPlus1 == pos)) frags.push_back(Frag{nKept, rec, pos});
                    std::memcpy(dst + nKept, line + i, run);
                    countRun(line + i, run, counts);
                    nKept += run;
                    haveLast = true; lastRec = rec; lastPosPlus1 = pos + run;
                }
                i = j;
            }
            seqLen += len;
        }
        seqLenEnd = seqLen;
    }
};

// The parser threads of one stream.  They live as long as the stream (a wave lasts a few milliseconds: threads created
// per wave would mostly still be waiting for a core when it ends); run(n, fn) executes fn(0..n-1) on them and on the
// caller and rethrows the first exception.
struct FastaStream::Pool {
    explicit Pool(unsigned helpers) { for (unsigned t = 0; t < helpers; t++) threads.emplace_back([this] { loop(); }); }
    ~Pool()
    {
        { std::lock_guard<std::mutex> l(m); stop = true; }
        wake.notify_all();
        for (auto& t : threads) t.join();
    }
    void run(size_t n, const std::function<void(size_t)>& fn)
    {
        if (n == 0) return;
        if (n == 1 || threads.empty()) { for (size_t i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> l(m);
            job = &fn; jobSize = n; nextIdx = 0; busy = threads.size(); err = nullptr; generation++;
        }
        wake.notify_all();
        work();
        std::unique_lock<std::mutex> l(m);
        idle.wait(l, [&] { return busy == 0; });
        job = nullptr;
        if (err) std::rethrow_exception(err);
    }
private:
    void work()
    {
        try { for (size_t i; (i = nextIdx.fetch_add(1)) < jobSize;) (*job)(i); }
        catch (...) { std::lock_guard<std::mutex> l(m); if (!err) err = std::current_exception(); }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m);
                wake.wait(l, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            work();
            std::lock_guard<std::mutex> l(m);
            if (--busy == 0) idle.notify_one();
        }
    }
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable wake, idle;
    const std::function<void(size_t)>* job = nullptr;
    size_t jobSize = 0, busy = 0;
    std::atomic<size_t> nextIdx{0};
    uint64_t generation = 0;
    bool stop = false;
    std::exception_ptr err;
};

FastaStream::FastaStream(const std::vector<std::string>& files, uint64_t maxFiltered)
    : files_(files), maxFiltered_(maxFiltered), pool_(new Pool(0)) {}

FastaStream::~FastaStream()
{
    if (map_) munmap(const_cast<char*>(map_), mapLen_);
    if (fd_ >= 0) close(fd_);
    std::free(buf_);
}

void FastaStream::setParallel(unsigned threads, size_t segmentBytes)
{
    threads_ = std::max(1u, threads);
    forcedSegment_ = segmentBytes;
    pool_.reset(new Pool(threads_ - 1));
}

bool FastaStream::openNext()
{
    if (map_) { munmap(const_cast<char*>(map_), mapLen_); map_ = nullptr; }
    if (fd_ >= 0) { close(fd_); fd_ = -1; }
    while (fileIdx_ < files_.size()) {
        const std::string& f = files_[fileIdx_++];
        fd_ = open(f.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("Could not open file: " + f);
        struct stat st;
        if (fstat(fd_, &st) != 0) throw std::runtime_error("Could not open file: " + f);
        mapLen_ = (size_t)st.st_size; mapPos_ = 0; posMidLine_ = false;
        if (mapLen_ == 0) { close(fd_); fd_ = -1; continue; }
        void* p = mmap(nullptr, mapLen_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (p == MAP_FAILED) throw std::runtime_error("Could not map file: " + f);
        madvise(p, mapLen_, MADV_SEQUENTIAL);
        map_ = static_cast<const char*>(p);
        return true;
    }
    return false;
}

void FastaStream::reserveBuf(size_t n)
{
    if (n <= bufCap_) return;
    const size_t cap = std::max(n, bufCap_ + bufCap_ / 2);
    char* p = static_cast<char*>(std::realloc(buf_, cap));
    if (!p) throw std::bad_alloc();
    buf_ = p; bufCap_ = cap;
}

// Serial part of a wave: give segment s its place in the stream.  p/n/midLine describe its bytes again because a
// segment that crosses the maxFiltered cut is re-parsed with the cut (the reference stops reading lines there, so
// headers behind the cut are not registered either).
void FastaStream::stitch(Segment& s, const char* p, size_t n, bool midLine)
{
    if (streamLen_ >= maxFiltered_) { eof_ = true; return; }
    if (s.nKept > maxFiltered_ - streamLen_) s.parse(p, n, midLine, maxFiltered_ - streamLen_);
    if (s.seqBeforeHeader && names_.empty()) throw std::runtime_error("Input file does not appear to be in fasta format\n");
    const uint64_t carrySeq = names_.empty() ? 0 : names_.size() - 1, carryLen = curSeqLen_, firstNew = names_.size();
    auto recOf = [&](uint32_t rec) { return rec == 0 ? carrySeq : firstNew + rec - 1; };
    auto posOf = [&](uint32_t rec, uint64_t pos) { return rec == 0 ? carryLen + pos : pos; };
    for (size_t k = 0; k < s.frags.size(); k++) {
        const auto& f = s.frags[k];
        const uint64_t seq = recOf(f.rec), pos = posOf(f.rec, f.pos);
        if (k == 0 && haveLast_ && lastSeq_ == seq && lastPosPlus1_ == pos) continue;     // continues the open fragment
        frags_.push_back(Fragment{streamLen_ + f.off, seq, pos});
    }
    if (s.haveLast) { haveLast_ = true; lastSeq_ = recOf(s.lastRec); lastPosPlus1_ = posOf(s.lastRec, s.lastPosPlus1); }
    curSeqLen_ = s.names.empty() ? carryLen + s.seqLenEnd : s.seqLenEnd;
    for (auto& nm : s.names) names_.push_back(std::move(nm));
    for (int i = 0; i < 4; i++) counts_[i] += s.counts[i];
    s.dstOff = bufHead_ + bufLen_;
    bufLen_ += s.nKept;
    streamLen_ += s.nKept;
}

void FastaStream::wave(uint64_t need)
{
    // FASTA carries 1/60 of line ends; a short wave is simply followed by another one
    const size_t left = mapLen_ - mapPos_;
    need = std::min<uint64_t>(need, 1ull << 40);
    const uint64_t needBytes = need + need / 32 + 4096;
    // four segments per thread: a thread that wakes up late (or never gets a core) costs a quarter of a share, not a whole one
    const size_t parts = threads_ > 1 ? 4 * (size_t)threads_ : 1;
    size_t segBytes = forcedSegment_;
    if (!segBytes) segBytes = (size_t)std::min<uint64_t>(std::max<uint64_t>(needBytes / parts + 1, 256 << 10), 64 << 20);
    size_t nSeg = forcedSegment_ ? threads_ : (size_t)std::min<uint64_t>(parts, (needBytes + segBytes - 1) / segBytes);
    nSeg = std::max<size_t>(1, std::min(nSeg, (left + segBytes - 1) / segBytes));

    // segment borders: line starts, except that a sequence line longer than 4 KB is cut in the middle
    struct Cut { size_t at; bool mid; };
    std::vector<Cut> cuts{{mapPos_, posMidLine_}};
    for (size_t k = 1; k <= nSeg; k++) {
        const Cut prev = cuts.back();
        const size_t x = std::min(mapLen_, mapPos_ + k * segBytes);
        if (x <= prev.at) continue;
        if (x == mapLen_) { cuts.push_back(Cut{x, false}); break; }
        const char* q = static_cast<const char*>(memrchr(map_ + prev.at, '\n', x - prev.at));
        if (q) {
            const size_t lineStart = (size_t)(q - map_) + 1;
            if (map_[lineStart] == '>' || x - lineStart <= 4096) { if (lineStart > prev.at) cuts.push_back(Cut{lineStart, false}); }
            else cuts.push_back(Cut{x, true});
        } else if (!prev.mid && map_[prev.at] == '>') {
            // inside a header line that began at the previous border: the border moves behind it
            const char* e = static_cast<const char*>(memchr(map_ + x, '\n', mapLen_ - x));
            cuts.push_back(Cut{e ? (size_t)(e - map_) + 1 : mapLen_, false});
        } else cuts.push_back(Cut{x, true});
    }
    const size_t n = cuts.size() - 1;
    if (n == 0) { mapPos_ = mapLen_; return; }
    while (segs_.size() < n) segs_.emplace_back(new Segment);
    pool_->run(n, [&](size_t k) {
        Segment& s = *segs_[k];
        s.src = map_ + cuts[k].at; s.srcLen = cuts[k + 1].at - cuts[k].at; s.midLine = cuts[k].mid;
        s.parse(s.src, s.srcLen, s.midLine, UINT64_MAX);
    });
    size_t used = 0, total = 0;
    for (; used < n && !eof_; used++) total += segs_[used]->nKept;      // upper bound before the cut
    // compact the buffer, make room, then place the segments
    if (bufHead_) { std::memmove(buf_, buf_ + bufHead_, bufLen_); bufHead_ = 0; }
    reserveBuf(bufLen_ + total);
    used = 0;
    for (; used < n; used++) {
        Segment& s = *segs_[used];
        stitch(s, s.src, s.srcLen, s.midLine);
        if (eof_) break;
    }
    pool_->run(used, [&](size_t k) {
        const Segment& s = *segs_[k];
        if (s.nKept) std::memcpy(buf_ + s.dstOff, s.chars.get(), s.nKept);
    });
    mapPos_ = cuts[n].at; posMidLine_ = cuts[n].mid;
}

bool FastaStream::fill(uint64_t want)
{
    while (bufLen_ < want && !eof_) {
        if (streamLen_ >= maxFiltered_) { eof_ = true; break; }
        if (!map_ || mapPos_ >= mapLen_) {
            if (!openNext()) { eof_ = true; break; }
        }
        wave(std::min<uint64_t>(want - bufLen_, maxFiltered_ - streamLen_));
    }
    return bufLen_ != 0;
}

bool FastaStream::next(uint64_t payload, uint64_t halo, Chunk& out)
{
    // drop what the previous chunk reported as payload; its halo becomes the head of this chunk
    if (pendingDrop_) {
        const uint64_t drop = std::min<uint64_t>(pendingDrop_, bufLen_);
        bufHead_ += drop; bufLen_ -= drop; bufStart_ += drop;
        pendingDrop_ = 0;
    }
    fill(payload + halo);
    // like the reference, a trailing chunk made only of the previous halo is still a chunk (its windows
    // were not reported yet); the stream ends when nothing is left at all (sequence.cpp:274-293)
    if (bufLen_ == 0) { out = Chunk(); return false; }
    out.chars = buf_ + bufHead_;
    out.nTotal = std::min<uint64_t>(bufLen_, payload + halo);
    out.nPayload = std::min<uint64_t>(out.nTotal, payload);
    out.streamStart = bufStart_;
    pendingDrop_ = out.nPayload;
    out.fragStarts.clear();
    out.frags.clear();
    auto it = std::upper_bound(frags_.begin(), frags_.end(), bufStart_,
                               [](uint64_t p, const Fragment& f) { return p < f.streamPos; });
    {   // the fragment that covers the first character of the chunk, re-based to chunk position 0
        const Fragment& f = *(it - 1);
        out.frags.push_back(Fragment{0, f.seqIdx, f.seqPos + (bufStart_ - f.streamPos)});
    }
    for (; it != frags_.end() && it->streamPos < bufStart_ + out.nTotal; ++it) {
        out.fragStarts.push_back(it->streamPos - bufStart_);
        out.frags.push_back(Fragment{it->streamPos - bufStart_, it->seqIdx, it->seqPos});
    }
    // fragments that lie wholly before the chunk are never needed again
    if (it - frags_.begin() > 4096) {
        auto keep = std::upper_bound(frags_.begin(), frags_.end(), bufStart_,
                                     [](uint64_t p, const Fragment& f) { return p < f.streamPos; }) - 1;
        frags_.erase(frags_.begin(), keep);
    }
    return true;
}

void FastaStream::copyChunk(const Chunk& c, char* dst)
{
    const size_t slice = 1 << 20, n = (size_t)((c.nTotal + slice - 1) / slice);
    pool_->run(n, [&](size_t k) { std::memcpy(dst + k * slice, c.chars + k * slice, std::min<size_t>(slice, c.nTotal - k * slice)); });
}

// ---- 2-bit packer (the device twin is csrc/pack.cuh: same codes, same mask, same padding) -------------------------------
// Sixteen characters per step as bit planes: with u = c & 0xDF, bit 2 of u is the high code bit and bit 1 ^ bit 2 the low one
// (A 0x41, C 0x43, G 0x47, T 0x54 -> 0 1 2 3, the row order of the matrix); a shift brings the wanted bit of every byte to
// bit 7, PMOVMSKB collects the sixteen of them, and one 64-bit spread interleaves the two planes into the code word.
namespace {
inline uint32_t interleave16(uint32_t lo, uint32_t hi)         // bit i of lo -> bit 2i, bit i of hi -> bit 2i + 1
{
    uint64_t x = (uint64_t)lo | ((uint64_t)hi << 32);
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x << 2)) & 0x3333333333333333ULL;
    x = (x | (x << 1)) & 0x5555555555555555ULL;
    return (uint32_t)x | ((uint32_t)(x >> 32) << 1);
}
#if defined(__SSE2__)
// 16 characters -> their code word; `zero` receives the 16 "contributes zero" bits
inline uint32_t pack16(const char* p, bool foldLower, uint32_t& zero)
{
    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
    const __m128i u = _mm_and_si128(x, _mm_set1_epi8((char)0xDF));
    // valid <=> the upper-cased byte is one of A C G T (a filtered stream holds nothing else; the test keeps the packer total:
    // any other byte, and the zero padding behind the block, becomes code 0 with its zero bit set)
    const __m128i v = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8(0x41)), _mm_cmpeq_epi8(u, _mm_set1_epi8(0x43))),
                                   _mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8(0x47)), _mm_cmpeq_epi8(u, _mm_set1_epi8(0x54))));
    const uint32_t valid = (uint32_t)_mm_movemask_epi8(v);
    const uint32_t hi = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(u, 5)) & valid;                                   // bit 2
    const uint32_t lo = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(_mm_xor_si128(u, _mm_srli_epi16(u, 1)), 6)) & valid;   // bit 1 ^ bit 2
    zero = ~valid & 0xFFFFu;
    if (!foldLower) zero |= (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(x, 2));                                      // bit 5: lower case
    return interleave16(lo, hi);
}
#else
// portable twin of the SSE2 routine (hosts without it, e.g. aarch64): the same rule, one character at a time
inline uint32_t pack16(const char* p, bool foldLower, uint32_t& zero)
{
    uint32_t code = 0; zero = 0;
    for (int i = 0; i < 16; i++) {
        const unsigned char x = (unsigned char)p[i], u = x & 0xDF;
        const bool valid = u == 0x41 || u == 0x43 || u == 0x47 || u == 0x54;
        if (valid) code |= (uint32_t)((((u >> 2) & 1u) << 1) | (((u >> 1) ^ (u >> 2)) & 1u)) << (2 * i);
        if (!valid || (!foldLower && (x & 0x20))) zero |= 1u << i;
    }
    return code;
}
#endif
#if defined(__x86_64__) && defined(__GNUC__)
// AVX2 + BMI2 twin (run-time dispatch): 32 characters per step, PDEP interleaves the two bit planes; twice the SSE2 rate
__attribute__((target("avx2,bmi2")))
uint64_t packWordsAvx2(const char* chars, uint64_t base, uint64_t i1, bool foldLower, uint32_t* codes2, uint32_t* zmask, bool& anyZero)
{
    const __m256i mDF = _mm256_set1_epi8((char)0xDF), cA = _mm256_set1_epi8(0x41), cC = _mm256_set1_epi8(0x43), cG = _mm256_set1_epi8(0x47),
                  cT = _mm256_set1_epi8(0x54);
    uint32_t any = 0;
    for (; base + 32 <= i1; base += 32) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(chars + base));
        const __m256i u = _mm256_and_si256(x, mDF);
        const __m256i v = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                          _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
        const uint32_t valid = (uint32_t)_mm256_movemask_epi8(v);
        const uint32_t hi = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(u, 5)) & valid;
        const uint32_t lo = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(_mm256_xor_si256(u, _mm256_srli_epi16(u, 1)), 6)) & valid;
        uint32_t zero = ~valid;
        if (!foldLower) zero |= (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 2));
        const uint64_t code = _pdep_u64(lo, 0x5555555555555555ULL) | _pdep_u64(hi, 0xAAAAAAAAAAAAAAAAULL);
        codes2[base / 16] = (uint32_t)code;
        codes2[base / 16 + 1] = (uint32_t)(code >> 32);
        zmask[base / 32] = zero;
        any |= zero;
    }
    anyZero |= any != 0;
    return base;
}
const bool kHaveAvx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && !getenv("BLAMM_B200_NO_AVX2");
// AVX-512BW twin: 64 characters per step, the byte tests deliver the bit planes directly as mask registers (no shift + MOVMSK)
__attribute__((target("avx512f,avx512bw,bmi2")))
uint64_t packWordsAvx512(const char* chars, uint64_t base, uint64_t i1, bool foldLower, uint32_t* codes2, uint32_t* zmask, bool& anyZero)
{
    const __m512i mDF = _mm512_set1_epi8((char)0xDF), cA = _mm512_set1_epi8(0x41), cC = _mm512_set1_epi8(0x43), cG = _mm512_set1_epi8(0x47),
                  cT = _mm512_set1_epi8(0x54), b1 = _mm512_set1_epi8(0x02), b2 = _mm512_set1_epi8(0x04), b5 = _mm512_set1_epi8(0x20);
    uint64_t any = 0;
    for (; base + 64 <= i1; base += 64) {
        const __m512i x = _mm512_loadu_si512(reinterpret_cast<const void*>(chars + base));
        const __m512i u = _mm512_and_si512(x, mDF);
        const uint64_t valid = _mm512_cmpeq_epi8_mask(u, cA) | _mm512_cmpeq_epi8_mask(u, cC) | _mm512_cmpeq_epi8_mask(u, cG) | _mm512_cmpeq_epi8_mask(u, cT);
        const uint64_t hi = _mm512_test_epi8_mask(u, b2) & valid;                                                   // bit 2
        const uint64_t lo = _mm512_test_epi8_mask(_mm512_xor_si512(u, _mm512_srli_epi16(u, 1)), b1) & valid;        // bit 1 ^ bit 2
        uint64_t zero = ~valid;
        if (!foldLower) zero |= _mm512_test_epi8_mask(x, b5);                                                       // bit 5: lower case
        const uint64_t c0 = _pdep_u64(lo & 0xFFFFFFFFu, 0x5555555555555555ULL) | _pdep_u64(hi & 0xFFFFFFFFu, 0xAAAAAAAAAAAAAAAAULL);
        const uint64_t c1 = _pdep_u64(lo >> 32, 0x5555555555555555ULL) | _pdep_u64(hi >> 32, 0xAAAAAAAAAAAAAAAAULL);
        codes2[base / 16] = (uint32_t)c0; codes2[base / 16 + 1] = (uint32_t)(c0 >> 32);
        codes2[base / 16 + 2] = (uint32_t)c1; codes2[base / 16 + 3] = (uint32_t)(c1 >> 32);
        zmask[base / 32] = (uint32_t)zero; zmask[base / 32 + 1] = (uint32_t)(zero >> 32);
        any |= zero;
    }
    anyZero |= any != 0;
    return base;
}
const bool kHaveAvx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2") && !getenv("BLAMM_B200_NO_AVX512") && !getenv("BLAMM_B200_NO_AVX2");
#endif
// characters [i0, i1) of `chars` (i0 a multiple of 32) -> their code and mask words; true if a live character contributes 0
bool packRange(const char* chars, uint64_t i0, uint64_t i1, bool foldLower, uint32_t* codes2, uint32_t* zmask)
{
    bool anyZero = false;
    uint64_t base = i0;
#if defined(__x86_64__) && defined(__GNUC__)
    if (kHaveAvx512) base = packWordsAvx512(chars, base, i1, foldLower, codes2, zmask, anyZero);
    if (kHaveAvx2) base = packWordsAvx2(chars, base, i1, foldLower, codes2, zmask, anyZero);
#endif
    for (; base + 32 <= i1; base += 32) {
        uint32_t z0, z1;
        codes2[base / 16] = pack16(chars + base, foldLower, z0);
        codes2[base / 16 + 1] = pack16(chars + base + 16, foldLower, z1);
        const uint32_t zm = z0 | (z1 << 16);
        zmask[base / 32] = zm;
        anyZero |= zm != 0;
    }
    if (base < i1) {                                            // ragged tail: zero padding (invalid: code 0, zero bit set)
        const uint64_t n = i1 - base;
        char tmp[32] = {0};
        std::memcpy(tmp, chars + base, (size_t)n);
        uint32_t z0, z1;
        codes2[base / 16] = pack16(tmp, foldLower, z0);
        const uint32_t c1 = pack16(tmp + 16, foldLower, z1);
        if (n > 16) codes2[base / 16 + 1] = c1;
        const uint32_t zm = z0 | (z1 << 16);
        zmask[base / 32] = zm;
        anyZero |= (zm & ((1u << n) - 1u)) != 0;
    }
    return anyZero;
}
} // namespace

bool packAscii(const char* chars, uint64_t n, bool foldLower, uint32_t* codes2, uint32_t* zmask)
{
    return packRange(chars, 0, n, foldLower, codes2, zmask);
}

bool FastaStream::packChunk(const Chunk& c, bool foldLower, uint32_t* codes2, uint32_t* zmask)
{
    const size_t slice = 1 << 20, n = (size_t)((c.nTotal + slice - 1) / slice);       // a multiple of 32: whole words per slice
    std::atomic<bool> anyZero{false};
    pool_->run(n, [&](size_t k) {
        if (packRange(c.chars, k * slice, std::min<uint64_t>(c.nTotal, (k + 1) * slice), foldLower, codes2, zmask)) anyZero = true;
    });
    return anyZero;
}

void FastaStream::locate(uint64_t streamPos, uint64_t& seqIdx, uint64_t& seqPos) const
{
    auto it = std::upper_bound(frags_.begin(), frags_.end(), streamPos,
                               [](uint64_t p, const Fragment& f) { return p < f.streamPos; });
    --it;
    seqIdx = it->seqIdx;
    seqPos = it->seqPos + (streamPos - it->streamPos);
}

} // namespace blamm
