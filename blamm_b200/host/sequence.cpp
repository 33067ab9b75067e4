// Species dictionary and the filtered sequence stream.
#include "host.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <set>
#include <sstream>

#include <fcntl.h>
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
// ---------------------------------------------------------------------------------------------------------
static const struct ValidTable {
    uint8_t v[256];
    ValidTable() { std::memset(v, 4, sizeof v);
                   v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; }
} kValid;

FastaStream::FastaStream(const std::vector<std::string>& files, uint64_t maxFiltered)
    : files_(files), maxFiltered_(maxFiltered) {}

FastaStream::~FastaStream()
{
    if (map_) munmap(const_cast<char*>(map_), mapLen_);
    if (fd_ >= 0) close(fd_);
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
        mapLen_ = (size_t)st.st_size; mapPos_ = 0;
        if (mapLen_ == 0) { close(fd_); fd_ = -1; continue; }
        void* p = mmap(nullptr, mapLen_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (p == MAP_FAILED) throw std::runtime_error("Could not map file: " + f);
        madvise(p, mapLen_, MADV_SEQUENTIAL);
        map_ = static_cast<const char*>(p);
        return true;
    }
    return false;
}

bool FastaStream::fill(uint64_t want)
{
    while (buf_.size() < want && !eof_) {
        if (streamLen_ >= maxFiltered_) { eof_ = true; break; }
        if (!map_ || mapPos_ >= mapLen_) {
            if (!openNext()) { eof_ = true; break; }
        }
        // one line
        const char* line = map_ + mapPos_;
        const char* nl = static_cast<const char*>(memchr(line, '\n', mapLen_ - mapPos_));
        const size_t len = nl ? (size_t)(nl - line) : mapLen_ - mapPos_;
        mapPos_ += len + (nl ? 1 : 0);
        if (len == 0) continue;
        if (line[0] == '>') {
            curSeqLen_ = 0;
            size_t b = 1;
            while (b < len && isspace((unsigned char)line[b])) b++;
            size_t e = b;
            while (e < len && !isspace((unsigned char)line[e])) e++;
            names_.emplace_back(line + b, e - b);
            continue;
        }
        if (names_.empty()) throw std::runtime_error("Input file does not appear to be in fasta format\n");
        const uint64_t seq = names_.size() - 1;
        size_t i = 0;
        while (i < len) {
            // skip a run of invalid characters
            while (i < len && kValid.v[(uint8_t)line[i]] == 4) i++;
            size_t j = i;
            while (j < len && kValid.v[(uint8_t)line[j]] != 4) j++;
            if (j == i) break;
            size_t run = j - i;
            if (streamLen_ + run > maxFiltered_) run = (size_t)(maxFiltered_ - streamLen_);
            if (run) {
                const uint64_t pos = curSeqLen_ + i;
                if (!(haveLast_ && lastSeq_ == seq && lastPosPlus1_ == pos))
                    frags_.push_back(Fragment{streamLen_, seq, pos});
                buf_.insert(buf_.end(), line + i, line + i + run);
                for (size_t k = i; k < i + run; k++) counts_[kValid.v[(uint8_t)line[k]]]++;
                streamLen_ += run;
                haveLast_ = true; lastSeq_ = seq; lastPosPlus1_ = pos + run;
            }
            i = j;
        }
        curSeqLen_ += len;
    }
    return !buf_.empty();
}

bool FastaStream::next(uint64_t payload, uint64_t halo, Chunk& out)
{
    // drop what the previous chunk reported as payload; its halo becomes the head of this chunk
    if (pendingDrop_) {
        const uint64_t drop = std::min<uint64_t>(pendingDrop_, buf_.size());
        buf_.erase(buf_.begin(), buf_.begin() + drop);
        bufStart_ += drop;
        pendingDrop_ = 0;
    }
    fill(payload + halo);
    // like the reference, a trailing chunk made only of the previous halo is still a chunk (its windows
    // were not reported yet); the stream ends when nothing is left at all (sequence.cpp:274-293)
    if (buf_.empty()) { out = Chunk(); return false; }
    out.chars = buf_.data();
    out.nTotal = std::min<uint64_t>(buf_.size(), payload + halo);
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

void FastaStream::locate(uint64_t streamPos, uint64_t& seqIdx, uint64_t& seqPos) const
{
    auto it = std::upper_bound(frags_.begin(), frags_.end(), streamPos,
                               [](uint64_t p, const Fragment& f) { return p < f.streamPos; });
    --it;
    seqIdx = it->seqIdx;
    seqPos = it->seqPos + (streamPos - it->streamPos);
}

} // namespace blamm
