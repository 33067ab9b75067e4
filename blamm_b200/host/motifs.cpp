// Settings, motifs, PWMs, thresholds and score histograms (host side of the scan path).
#include "host.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <fstream>
#include <iostream>
#include <sstream>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace blamm {

// ---------------------------------------------------------------------------------------------------------
// Settings  (keys and defaults: reference settings.cpp:34-74)
// ---------------------------------------------------------------------------------------------------------
Settings::Settings() { load("settings.cnf"); }
Settings::Settings(const std::string& path) { load(path); }

void Settings::load(const std::string& path)
{
    std::ifstream in(path);
    if (!in) return;
    defaultVal = false;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        if (key == "MATRIX_S_W") ls >> matrix_S_w;
        else if (key == "MATRIX_S_H") ls >> matrix_S_h;
        else if (key == "MATRIX_P_TILE_MIN_ZERO_AREA") ls >> matrix_P_tile_min_zero_area;
        else if (key == "PSEUDOCOUNT") ls >> pseudocount;
        else if (key == "FLUSHOUTPUT") ls >> flushOutput;
        else std::cerr << "WARNING: settings.cnf contains unknown key: " << key << std::endl;
    }
}

void Settings::print() const
{
    if (defaultVal) std::cout << "File settings.cnf not found, using default values" << std::endl;
    else std::cout << "Loaded configuration from file settings.cnf" << std::endl;
    std::cout << "  MATRIX_S_W = " << matrix_S_w << "; MATRIX_S_H = " << matrix_S_h
              << "; MATRIX_P_TILE_MIN_ZERO_AREA = " << matrix_P_tile_min_zero_area
              << "; PSEUDOCOUNT = " << pseudocount << "; FLUSHOUTPUT = " << flushOutput << "\n";
}

// ---------------------------------------------------------------------------------------------------------
// Motif
// ---------------------------------------------------------------------------------------------------------
static bool splitPermutationSuffix(const std::string& name, size_t& cut)
{
    // "<base>__<digits>" with a non-empty base (reference motif.h:256-281)
    ssize_t i = (ssize_t)name.size() - 1;
    while (i >= 0 && isdigit((unsigned char)name[i])) i--;
    if (i < 1) return false;
    if (name[i] != '_' || name[i - 1] != '_') return false;
    cut = (size_t)(i - 1);
    return true;
}

std::string Motif::baseName() const
{
    size_t cut;
    return splitPermutationSuffix(name, cut) ? name.substr(0, cut) : name;
}

bool Motif::isPermutation() const
{
    size_t cut;
    return splitPermutationSuffix(name, cut);
}

void Motif::computePWM(const std::array<uint64_t, 4>& bg, float pseudo)
{
    // background probabilities with pseudocount, complemented for a reverse-complement column
    float bgTot = (float)(bg[0] + bg[1] + bg[2] + bg[3]);
    bgTot += 4.0f * pseudo;
    std::array<float, 4> q;
    for (int b = 0; b < 4; b++) q[b] = ((float)bg[b] + pseudo) / bgTot;
    if (revComp) { std::swap(q[0], q[3]); std::swap(q[1], q[2]); }

    pwm.resize(pfm.size());
    for (size_t j = 0; j < pfm.size(); j++) {
        float tot = (float)(pfm[j][0] + pfm[j][1] + pfm[j][2] + pfm[j][3]);
        tot += 4.0f * pseudo;
        for (int b = 0; b < 4; b++) {
            const float p = ((float)pfm[j][b] + pseudo) / tot;
            pwm[j][b] = log2f(p / q[b]);
        }
    }
}

void Motif::reverseComplement()
{
    std::reverse(pfm.begin(), pfm.end());
    for (auto& r : pfm) { std::swap(r[0], r[3]); std::swap(r[1], r[2]); }
    std::reverse(pwm.begin(), pwm.end());
    for (auto& r : pwm) { std::swap(r[0], r[3]); std::swap(r[1], r[2]); }
    revComp = !revComp;
}

float Motif::maxScore() const
{
    float s = 0.0f;
    for (const auto& r : pwm) s += std::max(std::max(r[0], r[1]), std::max(r[2], r[3]));
    return s;
}

float Motif::minScore() const
{
    float s = 0.0f;
    for (const auto& r : pwm) s += std::min(std::min(r[0], r[1]), std::min(r[2], r[3]));
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// ScoreHistogram
// ---------------------------------------------------------------------------------------------------------
ScoreHistogram::ScoreHistogram(float mn, float mx, size_t bins) : counts(bins, 0), minScore(mn), maxScore(mx)
{
    width = (maxScore - minScore) / (float)bins;
}

static int binOf(float score, float mn, float width, size_t bins)
{
    int b = int((score - mn) / width);
    b = std::max(0, b);
    return std::min<int>((int)bins - 1, b);
}

void ScoreHistogram::setNumObservations(float score, uint64_t count) { counts[binOf(score, minScore, width, counts.size())] = count; }
void ScoreHistogram::addObservation(float score) { counts[binOf(score, minScore, width, counts.size())]++; }

float ScoreHistogram::scoreCutoff(float pvalue) const
{
    // walk down from the best bin until the requested fraction of observations is covered, then interpolate
    double total = 0.0;
    for (uint64_t c : counts) total += (double)c;
    double left = pvalue * total;
    for (ssize_t i = (ssize_t)counts.size() - 1; i >= 0; i--) {
        const double c = (double)counts[i];
        if (c < left) { left -= c; continue; }
        const double frac = left / c;
        const float edge = frac * i + (1.0 - frac) * (i + 1);
        return edge * width + minScore;
    }
    return maxScore;
}

// Histogram file -> counts (reference: ScoreHistogram::loadHistogram, motif.cpp:109-132): "bins min max" and then one "centre
// count" pair per bin, any white space between the numbers.  A `scan -pt` run loads one file per motif and group (21,600 for
// configs[2]), so the file is read with one call and parsed with std::from_chars (correctly rounded, like the stream extraction of
// the reference) instead of a stream.
void ScoreHistogram::load(const std::string& dir, const std::string& base)
{
    const std::string filename = dir + base + ".dat";
    std::string text;
    {
        FILE* f = fopen(filename.c_str(), "rb");
        if (!f) throw std::runtime_error("Error: cannot read file " + filename + ". Did you run the hist module?");
        char buf[1 << 16];
        for (size_t n; (n = fread(buf, 1, sizeof buf, f)) > 0;) text.append(buf, n);
        fclose(f);
    }
    const char* p = text.data();
    const char* const end = p + text.size();
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
    auto number = [&](auto& v) {
        while (p < end && blank(*p)) p++;
        if (p < end && *p == '+') p++;                              // (stream extraction accepts a leading plus sign)
        const auto r = std::from_chars(p, end, v);
        if (r.ec != std::errc()) throw std::runtime_error("Unexpected end-of-file reached");
        p = r.ptr;
    };
    uint64_t bins = 0;
    number(bins); number(minScore); number(maxScore);
    width = (maxScore - minScore) / (float)bins;
    counts.assign(bins, 0);
    for (uint64_t i = 0; i < bins; i++) {
        float centre;
        number(centre); number(counts[i]);
    }
}

// The two files of a histogram are assembled in memory and written with one call each (`hist` writes two files per motif and
// group: 43,200 files for configs[2]).  Number formatting: `ostream << x` in its default state is printf("%g", x), and
// std::to_chars(general, precision 6) is specified as exactly that conversion.  Contents as the reference's
// ScoreHistogram::writeGNUPlotFile (motif.cpp:71-107): bin centres are double expressions, the bounds floats.
namespace {
void putG(std::string& s, double v)
{
    char b[48];
    s.append(b, std::to_chars(b, b + sizeof b, v, std::chars_format::general, 6).ptr);
}
void putU(std::string& s, uint64_t v)
{
    char b[24];
    s.append(b, std::to_chars(b, b + sizeof b, v).ptr);
}
void spill(const std::string& filename, const std::string& bytes)
{
    FILE* f = fopen(filename.c_str(), "wb");
    const bool ok = f && fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    if (f && fclose(f) != 0) throw std::runtime_error("Error: cannot write to file " + filename);
    if (!ok) throw std::runtime_error("Error: cannot write to file " + filename);
}
}

void ScoreHistogram::writeGNUPlot(const std::string& dir, const std::string& base, const std::string& label) const
{
    std::string dat;
    dat.reserve(32 + 24 * counts.size());
    putU(dat, counts.size()); dat += '\t'; putG(dat, minScore); dat += '\t'; putG(dat, maxScore); dat += '\n';
    uint64_t tallest = 0;
    for (size_t i = 0; i < counts.size(); i++) {
        putG(dat, (0.5 + i) * width + minScore); dat += '\t'; putU(dat, counts[i]); dat += '\n';
        tallest = std::max(tallest, counts[i]);
    }
    spill(dir + base + ".dat", dat);

    // the plot script: y range = the tallest bin + 10 %, truncated (an integer scaled in double, as `size_t *= 1.1` does)
    const size_t yTop = (size_t)((double)(size_t)tallest * 1.1);
    std::string range;
    putG(range, minScore); range += ':'; putG(range, maxScore);
    std::string gnu = "set output \"" + base + ".ps\"\n"
                      "set key autotitle columnhead\n"
                      "set terminal postscript landscape\n"
                      "set terminal postscript noenhanced\n"
                      "set xrange [" + range + "]\n"
                      "set yrange [0:";
    putU(gnu, yTop);
    gnu += "]\nset xlabel 'PWM score'\nset ylabel 'count'\nplot \"" + base + ".dat\" using 1:2 title '" + label + "' with boxes\n";
    spill(dir + base + ".gnu", gnu);
}

// ---------------------------------------------------------------------------------------------------------
// MotifSet
// ---------------------------------------------------------------------------------------------------------
// JASPAR files, parsed from one in-memory copy with a cursor (10,000 PWMs for configs[3]).  The grammar is what the reference's
// stream extraction accepts (motif.cpp:377-407), quirks included:
//   * a record starts at the next white-space-delimited token, wherever it is (blank lines between records are skipped); its
//     first character is dropped whatever it is (normally '>'), the rest of that line is ignored;
//   * the next FOUR lines are the A, C, G, T rows, blank or not: two tokens are skipped ("A", "["), then unsigned integers are
//     taken until the first thing that is not one -- "12]" still yields 12, "[1" as second token swallows a count;
//   * a header token that ends the file yields no record.
// Rows shorter than the first one are padded with zeros (the reference reads past the end of the shorter row).
static bool isBlank(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

static void parseCounts(const char* p, const char* end, std::vector<uint64_t>& out)
{
    for (int skip = 0; skip < 2; skip++) {
        while (p < end && isBlank(*p)) p++;
        while (p < end && !isBlank(*p)) p++;
    }
    for (;;) {
        while (p < end && isBlank(*p)) p++;
        bool negative = false;
        if (p < end && (*p == '+' || *p == '-')) negative = *p++ == '-';
        uint64_t v = 0;
        const auto r = std::from_chars(p, end, v);
        if (r.ec != std::errc()) return;                   // no digits, or a count beyond 64 bits
        out.push_back(negative ? 0 - v : v);               // (stream extraction of an unsigned value negates modulo 2^64)
        p = r.ptr;
        if (p < end && !isBlank(*p)) return;               // "12]": the next extraction would start at the ']' and fail
    }
}

static void parseJaspar(const std::string& filename, std::vector<Motif>& out)
{
    std::string text;
    {
        std::ifstream in(filename, std::ios::binary);
        if (!in) throw std::runtime_error("Could not open file: " + filename);
        text.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    }
    const char* p = text.data();
    const char* const end = p + text.size();
    auto lineEnd = [end](const char* q) { const void* nl = memchr(q, '\n', (size_t)(end - q)); return nl ? static_cast<const char*>(nl) : end; };
    for (;;) {
        while (p < end && isBlank(*p)) p++;
        const char* tok = p;
        while (p < end && !isBlank(*p)) p++;
        if (p == end) break;                               // nothing left, or a header token that ends the file
        Motif m;
        m.name.assign(tok + 1, p);
        p = lineEnd(p);
        if (p < end) p++;
        std::vector<uint64_t> row[4];
        for (auto& r : row) {
            const char* e = lineEnd(p);
            parseCounts(p, e, r);
            p = e < end ? e + 1 : end;
        }
        m.pfm.reserve(row[0].size());
        for (size_t j = 0; j < row[0].size(); j++) {
            std::array<uint64_t, 4> col{};
            for (int b = 0; b < 4; b++) col[b] = j < row[b].size() ? row[b][j] : 0;
            m.pfm.push_back(col);
        }
        out.push_back(std::move(m));
    }
}

static void parseClusterBuster(const std::string& filename, std::vector<Motif>& out)
{
    // ">NAME" then one "A C G T" count row per position (reference motif.cpp:340-368)
    std::ifstream in(filename);
    if (!in) throw std::runtime_error("Could not open file: " + filename);
    std::string line;
    while (in.good()) {
        std::getline(in, line);
        if (line.empty()) continue;
        if (line.front() == '>') { Motif m; m.name = line.substr(1); out.push_back(m); continue; }
        if (out.empty()) throw std::runtime_error("Incorrect motif file format: " + filename);
        std::istringstream ls(line);
        size_t a = 0, c = 0, g = 0, t = 0;
        ls >> a >> c >> g >> t;
        out.back().pfm.push_back({a, c, g, t});
    }
}

void MotifSet::load(const std::string& filename, bool loadPermutations)
{
    std::vector<Motif> all;
    const bool jaspar = filename.size() > 7 && filename.compare(filename.size() - 7, 7, ".jaspar") == 0;
    if (jaspar) parseJaspar(filename, all); else parseClusterBuster(filename, all);
    // ascending length; same comparator and the same std::sort as the reference (motif.cpp:436, motif.h:318),
    // so ties land in the same (implementation-defined) order and column indices agree with the reference's
    std::sort(all.begin(), all.end(), [](const Motif& a, const Motif& b) { return a.size() < b.size(); });
    for (auto& m : all)
        if (loadPermutations || !m.isPermutation()) motifs.push_back(std::move(m));
}

void MotifSet::addReverseComplements()
{
    std::vector<Motif> both;
    both.reserve(2 * motifs.size());
    for (auto& m : motifs) {
        both.push_back(m);
        Motif r = m;
        r.reverseComplement();
        both.push_back(std::move(r));
    }
    motifs.swap(both);
}

size_t MotifSet::maxLen() const
{
    size_t n = 0;
    for (const auto& m : motifs) n = std::max(n, m.size());
    return n;
}

void MotifSet::generateMatrix(const std::array<uint64_t, 4>& bgCounts, float pseudo)
{
    for (auto& m : motifs) m.computePWM(bgCounts, pseudo);
    const size_t ld = 4 * maxLen();
    P_.assign(ld * motifs.size(), 0.0f);
    for (size_t c = 0; c < motifs.size(); c++)
        for (size_t j = 0; j < motifs[c].size(); j++)
            for (int b = 0; b < 4; b++) P_[c * ld + 4 * j + b] = motifs[c].pwm[j][b];
}

std::vector<int32_t> MotifSet::colLen() const
{
    std::vector<int32_t> v;
    for (const auto& m : motifs) v.push_back((int32_t)m.size());
    return v;
}

std::vector<float> MotifSet::colThr() const
{
    std::vector<float> v;
    for (const auto& m : motifs) v.push_back(m.threshold);
    return v;
}

void MotifSet::theoreticalHistogram(const Motif& m, const std::array<float, 4>& bg, size_t numBins, uint64_t maxLength,
                                    ScoreHistogram& hist)
{
    // Discretise the PWM to integers so that the score range maps onto numBins steps, convolve the
    // per-position score distributions under the background, and store maxLength * pdf per bin
    // (reference motif.cpp:151-192 + hist.cpp:162-175; maps walked in ascending key order there, dense
    // arrays walked in ascending index order here -- same float additions in the same order).
    const size_t L = m.size();
    const float minS = m.minScore(), maxS = m.maxScore();
    const float a = (float)numBins / (maxS - minS);
    const float b = -a * minS / (float)L;
    std::vector<std::array<int, 4>> w(L);
    long lo = 0, hi = 0, runLo = 0, runHi = 0;
    for (size_t p = 0; p < L; p++) {
        for (int k = 0; k < 4; k++) w[p][k] = (int)std::round(a * m.pwm[p][k] + b);
        runLo += *std::min_element(w[p].begin(), w[p].end());
        runHi += *std::max_element(w[p].begin(), w[p].end());
        lo = std::min(lo, runLo); hi = std::max(hi, runHi);
    }
    const long span = hi - lo + 1;
    std::vector<float> cur(span, 0.0f), nxt(span, 0.0f);
    std::vector<char> curSet(span, 0), nxtSet(span, 0);
    for (int k = 0; k < 4; k++) { cur[w[0][k] - lo] += bg[k]; curSet[w[0][k] - lo] = 1; }
    for (size_t p = 1; p < L; p++) {
        std::fill(nxt.begin(), nxt.end(), 0.0f); std::fill(nxtSet.begin(), nxtSet.end(), 0);
        for (long s = 0; s < span; s++) {
            if (!curSet[s]) continue;
            for (int k = 0; k < 4; k++) { nxt[s + w[p][k]] += cur[s] * bg[k]; nxtSet[s + w[p][k]] = 1; }
        }
        cur.swap(nxt); curSet.swap(nxtSet);
    }
    for (long s = 0; s < span; s++) {
        if (!curSet[s]) continue;
        const float score = ((float)(s + lo) - L * b) / a;
        hist.setNumObservations(score, (uint64_t)(maxLength * cur[s]));
    }
}

// "%g" (precision 6) of a float, the format of `ostream << float` in the reference's writer (pwmscan.cpp:94).
// Fast path for 1e-4 <= |v| < 1e6, i.e. the fixed notation of %g: with X = floor(log10 |v|), y = |v| * 10^(5 - X) is EXACT in
// double arithmetic (a float has 24 significant bits and 10^k = 2^k * 5^k with 5^9 < 2^21, so the product needs at most 45
// bits), hence nearbyint(y) -- round to nearest, ties to even -- is the correctly rounded 6-digit decimal that printf
// produces from the exact binary value; the digits are then laid out as %f with precision 5 - X and the trailing zeros are
// dropped.  The comparisons against the double constants 1e-4 .. 1e-1 are exact for float inputs (no float lies between
// such a constant and the decimal it approximates).  Everything else (zero, exponent notation, non-finite) goes to
// std::to_chars / snprintf.  tests/test_host.py checks the function against printf on random bit patterns, ties and edges.
namespace {
// "000" .. "999", four bytes apart: the six significant digits come out of two table reads
struct Digits3 {
    char d[1000][4];
    Digits3() { for (int i = 0; i < 1000; i++) { d[i][0] = (char)('0' + i / 100); d[i][1] = (char)('0' + i / 10 % 10); d[i][2] = (char)('0' + i % 10); d[i][3] = 0; } }
};
const Digits3 kDigits3;

// y in [1e5, 1e6], exact in double: round to nearest, ties to even (what printf does with the exact binary value).
// x86-64: CVTSD2SI under the default MXCSR rounding mode; elsewhere nearbyint under the default rounding mode.
inline uint32_t roundEven(double y)
{
#if defined(__SSE2__)
    return (uint32_t)_mm_cvtsd_si32(_mm_set_sd(y));
#else
    return (uint32_t)std::nearbyint(y);
#endif
}
}

int formatScore(char* dst, float v)
{
    static const double kPow10[10] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};
    const double a = std::fabs((double)v);
    if (a >= 1e-4 && a < 1e6) {
        int X = a >= 1e2 ? (a >= 1e4 ? (a >= 1e5 ? 5 : 4) : (a >= 1e3 ? 3 : 2))
                         : (a >= 1e0 ? (a >= 1e1 ? 1 : 0) : (a >= 1e-2 ? (a >= 1e-1 ? -1 : -2) : (a >= 1e-3 ? -3 : -4)));
        uint32_t r = roundEven(a * kPow10[5 - X]);                          // 100000 .. 1000000
        if (r >= 1000000u) { r = 100000u; X++; }                            // 999999.5 -> 1.00000 x 10^(X+1)
        if (X <= 5) {
            const uint32_t hi = r / 1000u, lo = r - 1000u * hi;
            char dig[16] = {};                                               // (16: the 8-byte copies below start at up to dig + 6)
            std::memcpy(dig, kDigits3.d[hi], 4);
            std::memcpy(dig + 3, kDigits3.d[lo], 4);
            int nd = 6;
            while (nd > 1 && dig[nd - 1] == '0') nd--;                      // significant digits left after dropping trailing zeros
            char* p = dst;
            if (std::signbit(v)) *p++ = '-';
            if (X >= 0) {
                // all six digits, then the fraction moved one place right behind the point (the caller's buffer has room: a score
                // takes at most 16 characters and every writer leaves 32); zeros inside the integer part are digits, not trailing zeros
                std::memcpy(p, dig, 8);
                if (nd <= X + 1) return (int)(p - dst) + X + 1;
                p[X + 1] = '.';
                std::memcpy(p + X + 2, dig + X + 1, 8);
                return (int)(p - dst) + nd + 1;
            }
            *p++ = '0'; *p++ = '.';
            for (int i = -1; i > X; i--) *p++ = '0';
            std::memcpy(p, dig, 8);
            return (int)(p - dst) + nd;
        }
    }
    if (!std::isfinite(v)) return snprintf(dst, 32, "%g", (double)v);
    return (int)(std::to_chars(dst, dst + 32, (double)v, std::chars_format::general, 6).ptr - dst);
}


} // namespace blamm
