// blamm-b200: command line front end (dict / hist / scan) with the reference's flags, inputs and outputs.
// The scan module drives libb200scan.so through its C ABI (include/b200scan.h); there is no CPU scoring path.
#include "host.h"
#include "../../include/b200scan.h"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <map>
#include <mutex>
#include <random>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <csignal>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/vfs.h>
#include <unistd.h>

using namespace std;

namespace blamm {

// operator<<(ostream&, float) of the reference (pwmscan.cpp:93) == printf("%g").  std::to_chars in general format with
// precision 6 is specified to give exactly that conversion (checked against snprintf on 4e7 floats) at a fifth of the cost.
namespace {
struct PhaseTimer {            // BLAMM_B200_TIMING=1: wall-clock seconds per phase on stderr (sums over threads where noted)
    bool on = getenv("BLAMM_B200_TIMING") != nullptr, quiet = false;
    std::mutex m; std::vector<std::pair<std::string, double>> acc;
    void add(const char* what, double s) {
        if (!on) return;
        std::lock_guard<std::mutex> l(m);
        for (auto& a : acc) if (a.first == what) { a.second += s; return; }
        acc.emplace_back(what, s);
    }
    void report() { if (on && !quiet) for (auto& a : acc) fprintf(stderr, "[timing] %-28s %8.3f s\n", a.first.c_str(), a.second); }
} gTimer;
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// `scan --stats FILE`: machine-readable account of the run (phases of gTimer + per-GPU sums of the b200scan_timing records)
struct DeviceStats { uint64_t chunks = 0, chars = 0, hits = 0, candidates = 0; double h2d = 0, pack = 0, score = 0, rescore = 0, order = 0, d2h = 0; };
struct RunStats {
    std::string file; std::mutex m; std::vector<DeviceStats> dev;
    void add(size_t d, const b200scan_timing& t, uint64_t chars) {
        std::lock_guard<std::mutex> l(m);
        if (dev.size() <= d) dev.resize(d + 1);
        DeviceStats& s = dev[d];
        s.chunks++; s.chars += chars; s.hits += t.n_hits; s.candidates += t.n_candidates;
        s.h2d += t.h2d_ms; s.pack += t.pack_ms; s.score += t.score_ms; s.rescore += t.rescore_ms; s.order += t.order_ms; s.d2h += t.d2h_ms;
    }
} gStats;
uint64_t gStatsColumns = 0, gStatsMatches = 0;
// FASTA parser threads: -t where the module has it, else the host's cores; BLAMM_B200_INGEST_THREADS overrides (1 = serial)
unsigned ingestThreads(size_t wanted = 0)
{
    if (const char* e = getenv("BLAMM_B200_INGEST_THREADS")) return (unsigned)std::max(1, atoi(e));
    const size_t t = wanted ? wanted : std::thread::hardware_concurrency();
    return (unsigned)std::min<size_t>(std::max<size_t>(t, 1), 64);
}
}

// =========================================================================================================
// dict  (reference dict.cpp:53-115)
// =========================================================================================================
static void dictUsage()
{
    cout << "Usage: blamm dict [options] sequences.input\n"
            "Goal: compute nucleotide frequencies for the input sequences and generate a sequence dictionary file\n\n"
            " [options]\n  -h\t--help\t\tdisplay help message\n\n"
            " File \"sequences.input\" contains a list of input fasta files in the following format:\n"
            "   speciesID_1\tspecies1_sequences.fasta\n   speciesID_2\tspecies2_sequences.fasta\n"
            "   speciesID_3\tspecies2_sequences.fasta\n   ...\n"
            " where speciesID_x is a user-defined identifier per species\n\n"
            " Example:\n  blamm dict sequences.input\n\n";
}

int runDict(int argc, char** argv)
{
    if (argc < 3) { dictUsage(); return EXIT_FAILURE; }
    for (int i = 2; i < argc - 1; i++) {
        string arg(argv[i]);
        dictUsage();
        return (arg == "-h" || arg == "--help") ? EXIT_SUCCESS : EXIT_FAILURE;
    }
    cout << "Welcome to blamm -- dictionary model" << endl;
    const string manifest(argv[argc - 1]);
    ifstream in(manifest);
    if (!in) throw runtime_error("Could not open file: " + manifest);
    SpeciesSet set;
    string line;
    while (getline(in, line)) {
        if (line.empty()) continue;
        string group, fasta;
        istringstream ls(line);
        ls >> group >> fasta;
        if (group.empty() || fasta.empty())
            throw runtime_error("File " + manifest + " has incorrect format.\nRefer to the documention for more information.");
        if (!ifstream(fasta)) throw runtime_error("Could not open fasta file: " + fasta);
        set.addFile(group, fasta);
    }
    for (auto& s : set.species) {
        cout << "Computing nucleotide composition for " << s.name << " ...";
        cout.flush();
        FastaStream fs(s.files);
        fs.setParallel(ingestThreads());
        FastaStream::Chunk c;
        while (fs.next(64 << 20, 0, c)) {}
        s.nuclCounts = fs.counts();
        s.totSeqLen = s.nuclCounts[0] + s.nuclCounts[1] + s.nuclCounts[2] + s.nuclCounts[3];
        s.seqNames = fs.seqNames();
        cout << "\n";
    }
    set.writeDict(manifest + ".dict");
    cout << "Done!  Dictionary written to " << manifest + ".dict" << endl;
    return EXIT_SUCCESS;
}

// =========================================================================================================
// hist  (reference hist.cpp:177-266).  Theoretical spectra are computed on the host (cheap); the empirical mode `-e`
//        scores the first -l characters of every group on the GPU (b200scan_hist_*), never on the CPU.
// =========================================================================================================
static void histUsage()
{
    cout << "Usage: blamm hist [options] motifs.input sequences.input\n"
            "Goal: compute PWM score histograms\n\n"
            " [options]\n  -h\t--help\t\tdisplay help message\n\n"
            " [options arg]\n"
            "  -l\t--length\tmaximum sequence length to analyze (default = 10000000)\n"
            "  -b\t--numbins\tnumber of bins per histogram (default = 250)\n"
            "  -t\t--numthreads\tset the number of parallel threads [default = #cores]\n"
            "  -e\t--empirical\tscore the sequences themselves (on the GPUs) instead of computing the theoretical spectrum\n"
            "  -g\t--gpus\t\tnumber of GPUs for -e [default = all]\n\n"
            " [file_options]\n  -H\t--histdir\toutput directory for the histogram file(s) [default = .]\n\n"
            " Example:\n  blamm hist motifs.input sequences.input\n\n";
}

int runHist(int argc, char** argv)
{
    if (argc < 4) { histUsage(); return EXIT_FAILURE; }
    uint64_t maxLength = 10000000; size_t numBins = 250; bool empirical = false; string histdir;
    size_t numThreads = thread::hardware_concurrency();
    int gpusWanted = 0;
    for (int i = 2; i < argc - 2; i++) {
        string arg(argv[i]);
        const bool hasVal = i + 1 < argc - 2;
        if (arg == "-h" || arg == "--help") { histUsage(); return EXIT_SUCCESS; }
        else if ((arg == "-l" || arg == "--length") && hasVal) maxLength = atoll(argv[++i]);
        else if ((arg == "-b" || arg == "--numbins") && hasVal) { numBins = atoll(argv[++i]); if (numBins < 2) numBins = 2; }
        else if (arg == "-e" || arg == "--empirical") empirical = true;
        else if ((arg == "-t" || arg == "--numthreads") && hasVal) numThreads = (size_t)max(1, atoi(argv[++i]));
        else if ((arg == "-g" || arg == "--gpus") && hasVal) gpusWanted = atoi(argv[++i]);
        else if ((arg == "-H" || arg == "--histdir") && hasVal) { histdir = argv[++i]; if (histdir.back() != '/') histdir.push_back('/'); }
        else { histUsage(); return EXIT_FAILURE; }
    }
    // one task per motif on -t threads: the spectrum DP and the two files of a motif are independent of every other motif
    // (the reference computes and writes them one after the other, hist.cpp:162-175)
    auto forEachMotif = [&](size_t n, const function<void(size_t)>& fn) {
        const size_t T = max<size_t>(1, min<size_t>(min<size_t>(numThreads, 64), n));
        atomic<size_t> next(0);
        vector<string> err(T);
        auto work = [&](size_t t) {
            try { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); } catch (const exception& e) { err[t] = e.what(); next = n; }
        };
        vector<thread> pool;
        for (size_t t = 1; t < T; t++) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        for (const auto& e : err) if (!e.empty()) throw runtime_error(e);
    };
    cout << "Welcome to blamm -- histogram module" << endl;
    Settings settings;
    SpeciesSet sc;
    sc.loadDict(string(argv[argc - 1]) + ".dict");
    MotifSet mc;
    mc.load(argv[argc - 2], false);
    cout << "Loaded " << mc.motifs.size() << " motifs from disk";
    cout << "\nMaximum motif size: " << mc.maxLen() << endl;
    if (mc.motifs.empty()) throw runtime_error("No motifs found in " + string(argv[argc - 2]));
    // `hist -e`: one context per GPU (-g, default all); the chunks of a group are dealt to the devices, every device keeps its own
    // 64-bit bin counters, the host adds them up (integer sums: the result does not depend on the deal).  The reference runs
    // histThread on -t host threads over 250,000-character blocks (hist.cpp:95-160).
    vector<b200scan_ctx*> ctxs;
    const uint64_t halo = mc.maxLen() - 1;
    const uint64_t chunk = std::min<uint64_t>(maxLength, 32ull << 20);
    auto destroyAll = [&] { for (auto c : ctxs) if (c) b200scan_destroy(c); ctxs.clear(); };
    if (empirical) {
        if (mc.maxLen() > B200SCAN_MAX_MOTIF_LEN)
            throw runtime_error("Motifs longer than " + to_string(B200SCAN_MAX_MOTIF_LEN) + " positions are not supported by this build");
        int nDev = b200scan_device_count();
        if (nDev == 0) throw runtime_error("CUDA error: no sm_100 devices found. Aborting...");
        if (gpusWanted > 0) nDev = min(nDev, gpusWanted);
        ctxs.assign((size_t)nDev, nullptr);
        vector<future<string>> made;
        for (int d = 0; d < nDev; d++)
            made.push_back(async(launch::async, [&ctxs, d, chunk, halo]() -> string {
                return b200scan_create(&ctxs[(size_t)d], d, chunk + halo + 64, 1024) == B200SCAN_OK ? string() : string(b200scan_last_error(nullptr));
            }));
        string err;
        const double tCreate = now();
        for (auto& m : made) { const string e = m.get(); if (!e.empty() && err.empty()) err = e; }
        gTimer.add("hist -e: CUDA start-up (contexts of all GPUs)", now() - tCreate);
        if (!err.empty()) { destroyAll(); throw runtime_error("CUDA error: " + err); }
    }
    future<void> pendingWrite;
    for (const auto& sp : sc.species) {
        double tPhase = now();                      // BLAMM_B200_TIMING=1: the phases of the module on stderr
        cout << "Generating histograms for species: " << sp.name;
        sp.printNuclProb(settings.pseudocount);
        mc.generateMatrix(sp.nuclCounts, settings.pseudocount);
        const auto bg = sp.nuclProb(settings.pseudocount);
        vector<ScoreHistogram> hists;
        for (const auto& m : mc.motifs) hists.emplace_back(m.minScore(), m.maxScore(), numBins);
        gTimer.add("hist: matrix P per group", now() - tPhase); tPhase = now();
        if (!empirical) {
            forEachMotif(mc.motifs.size(), [&](size_t i) { MotifSet::theoreticalHistogram(mc.motifs[i], bg, numBins, maxLength, hists[i]); });
            gTimer.add("hist: theoretical spectra", now() - tPhase); tPhase = now();
        } else {
            // reference: FastaBatch(filenames, maxLength) + histThread (hist.cpp:95-160)
            const auto len = mc.colLen();
            vector<float> thr(len.size(), 0.0f), mn, mx;
            for (const auto& h : hists) { mn.push_back(h.minScore); mx.push_back(h.maxScore); }
            struct HistJob { unique_ptr<char[]> chars; vector<uint64_t> fragStarts; uint64_t nTotal = 0, nPayload = 0; };
            mutex qm; condition_variable qcv; deque<unique_ptr<HistJob>> q; bool done = false; string failure;
            auto worker = [&](b200scan_ctx* ctx) {
                auto fail = [&](const string& e) { lock_guard<mutex> l(qm); if (failure.empty()) failure = e; qcv.notify_all(); };
                double t0 = now();
                if (b200scan_set_motifs(ctx, mc.P().data(), mc.ldp(), (int32_t)len.size(), len.data(), thr.data()) != B200SCAN_OK ||
                    b200scan_hist_begin(ctx, mn.data(), mx.data(), (uint32_t)numBins) != B200SCAN_OK) { fail(b200scan_last_error(ctx)); return; }
                gTimer.add("hist -e: set_motifs + hist_begin (summed over GPUs)", now() - t0);
                for (;;) {
                    unique_ptr<HistJob> job;
                    t0 = now();
                    {
                        unique_lock<mutex> l(qm);
                        qcv.wait(l, [&] { return !q.empty() || done || !failure.empty(); });
                        if (!failure.empty() || q.empty()) return;
                        job = std::move(q.front()); q.pop_front();
                        qcv.notify_all();
                    }
                    gTimer.add("hist -e: workers wait for chunks (summed over GPUs)", now() - t0); t0 = now();
                    if (b200scan_hist_block_ascii(ctx, job->chars.get(), job->nTotal, job->nPayload, job->fragStarts.data(), job->fragStarts.size(),
                                                  B200SCAN_LOWER_ZERO) != B200SCAN_OK) { fail(b200scan_last_error(ctx)); return; }
                    gTimer.add("hist -e: hist_block (previous block's kernels + staging, summed over GPUs)", now() - t0);
                }
            };
            vector<thread> workers;
            for (auto c : ctxs) workers.emplace_back(worker, c);
            try {
                FastaStream fs(sp.files, maxLength);
                fs.setParallel(ingestThreads(numThreads));
                FastaStream::Chunk c;
                // at least two chunks per GPU and group, so that every device works on every group (not below 4 Mi characters)
                const uint64_t groupLen = std::min<uint64_t>(maxLength, sp.totSeqLen ? sp.totSeqLen : maxLength);
                uint64_t chunkSp = std::min<uint64_t>(chunk, std::max<uint64_t>(4ull << 20, groupLen / (2 * ctxs.size()) + 1));
                if (const char* e = getenv("BLAMM_B200_CHUNK")) chunkSp = std::min<uint64_t>(chunk, std::max<uint64_t>(strtoull(e, nullptr, 10), 1024));
                while (fs.next(chunkSp, halo, c)) {
                    unique_ptr<HistJob> job(new HistJob);
                    job->chars.reset(new char[c.nTotal]);
                    fs.copyChunk(c, job->chars.get());
                    job->fragStarts = c.fragStarts; job->nTotal = c.nTotal; job->nPayload = c.nPayload;
                    unique_lock<mutex> l(qm);
                    qcv.wait(l, [&] { return q.size() < ctxs.size() + 1 || !failure.empty(); });
                    if (!failure.empty()) break;
                    q.push_back(std::move(job));
                    qcv.notify_all();
                    cout << "."; cout.flush();
                }
            } catch (...) {
                { lock_guard<mutex> l(qm); done = true; if (failure.empty()) failure = "input error"; }
                qcv.notify_all();
                for (auto& w : workers) w.join();
                destroyAll();
                throw;
            }
            gTimer.add("hist -e: read FASTA + deal chunks (reader, incl. waiting for queue room)", now() - tPhase); tPhase = now();
            { lock_guard<mutex> l(qm); done = true; }
            qcv.notify_all();
            for (auto& w : workers) w.join();
            cout << endl;
            if (!failure.empty()) { destroyAll(); throw runtime_error("CUDA error: " + failure); }
            vector<uint64_t> counts(len.size() * numBins), total(len.size() * numBins, 0);
            for (auto c : ctxs) {
                if (b200scan_hist_read(c, counts.data(), counts.size()) != B200SCAN_OK) { const string e = b200scan_last_error(c); destroyAll(); throw runtime_error("CUDA error: " + e); }
                for (size_t i = 0; i < counts.size(); i++) total[i] += counts[i];
            }
            for (size_t i = 0; i < hists.size(); i++)
                for (size_t b = 0; b < numBins; b++) hists[i].counts[b] = total[i * numBins + b];
            gTimer.add("hist -e: last kernels + counters of all GPUs", now() - tPhase); tPhase = now();
        }
        // the two files per motif are written in the background while the next group is read and scored
        if (pendingWrite.valid()) pendingWrite.get();
        gTimer.add("hist: wait for the previous group's files", now() - tPhase);
        vector<string> names;
        for (const auto& m : mc.motifs) names.push_back(m.name);
        pendingWrite = async(launch::async, [&forEachMotif, &histdir, group = sp.name, names = std::move(names), hs = std::move(hists)] {
            const double t0 = now();
            forEachMotif(hs.size(), [&](size_t i) { hs[i].writeGNUPlot(histdir, "hist_" + group + "_" + names[i], names[i] + " (" + group + ")"); });
            gTimer.add("hist: write the histogram files (background)", now() - t0);
        });
    }
    if (pendingWrite.valid()) {
        const double t0 = now();
        try { pendingWrite.get(); } catch (...) { destroyAll(); throw; }
        gTimer.add("hist: wait for the last group's files", now() - t0);
    }
    destroyAll();
    gTimer.report();
    return EXIT_SUCCESS;
}

// =========================================================================================================
// scan
// =========================================================================================================
static void scanUsage()
{
    cout << "Usage: blamm scan [options] motifs.input sequences.input\n"
            "Goal: find PWM matches in sequences\n\n"
            " [options]\n"
            "  -h\t--help\t\tdisplay help message\n"
            "  -s\t--simple\tscore lower-case nucleotides like upper case (the reference's simple-scan semantics)\n"
            "  -c\t--cuda\t\taccepted for compatibility (this build always scans on the GPU)\n"
            "  -rc\t--revcompl\talso search the reverse strand for occurrences\n\n"
            " [options arg]\n"
            "  -at\t--absthreshold\tset the minimal absolute score for a motif occurrence\n"
            "  -rt\t--relthreshold\tset the minimal relative score [0..1] for a motif occurrence (default = 0.95)\n"
            "  -pt\t--pthreshold\tcompute the motif score threshold from p-value [0..1] (default = 1E-4)\n"
            "  -t\t--numthreads\tset the number of host formatting threads [default = #cores]\n"
            "  -g\t--gpus\t\tnumber of GPUs to use [default = all]\n"
            "  -e\t--engine\tauto | tensor | gather [default = auto]\n\n"
            " [file_options]\n"
            "  -H\t--histdir\tdirectory where the histogram file(s) are stored [default = .]\n"
            "  -o\t--output\tfilename for the motif occurrences [default = occurences.txt]\n"
            "  \t--stats\tfilename for a JSON account of the run (phases, per-GPU kernel and transfer times)\n\n"
            " File \"motifs.input\" should contain the motifs in Jaspar format\n"
            " File \"sequences.input\" should contain a list of input fasta files (see blamm dict)\n\n"
            " Example:\n  blamm scan -o occurences.txt motifs.input sequences.input\n\n";
}

namespace {

// A pool of host threads shared by everything that works on hit lists (formatting, file writes): however many GPUs feed the
// run, the occurrence writer uses all -t threads.  parallel(n, fn) runs fn(0 .. n-1) on the pool AND on the calling thread and
// returns when all are done; calls from several threads queue up behind one another.
class WorkPool {
public:
    explicit WorkPool(size_t threads) { for (size_t t = 0; t + 1 < threads; t++) workers.emplace_back([this] { loop(); }); }
    ~WorkPool() { { lock_guard<mutex> l(m); stop = true; } cv.notify_all(); for (auto& t : workers) t.join(); }
    size_t size() const { return workers.size() + 1; }          // threads a parallel() call can occupy (the pool's + the caller)
    void parallel(size_t n, const function<void(size_t)>& fn)
    {
        if (n == 0) return;
        if (n == 1 || workers.empty()) { for (size_t i = 0; i < n; i++) fn(i); return; }
        auto t = make_shared<Task>();
        t->fn = &fn; t->n = n;
        { lock_guard<mutex> l(m); q.push_back(t); }
        cv.notify_all();
        for (size_t i; (i = t->next.fetch_add(1)) < n;) { run(*t, i); t->done.fetch_add(1); }
        {
            unique_lock<mutex> l(m);
            cvDone.wait(l, [&] { return t->done.load() == n; });
            for (auto it = q.begin(); it != q.end(); ++it) if (*it == t) { q.erase(it); break; }
        }
        if (t->err) rethrow_exception(t->err);            // an exception of any item reaches the caller, not std::terminate
    }
private:
    struct Task { const function<void(size_t)>* fn = nullptr; size_t n = 0; atomic<size_t> next{0}, done{0}; mutex em; exception_ptr err; };
    static void run(Task& t, size_t i)
    {
        try { (*t.fn)(i); }
        catch (...) { lock_guard<mutex> l(t.em); if (!t.err) t.err = current_exception(); }
    }
    void loop()
    {
        for (;;) {
            shared_ptr<Task> t;
            {
                unique_lock<mutex> l(m);
                cv.wait(l, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                t = q.front();
            }
            const size_t i = t->next.fetch_add(1);
            if (i >= t->n) {                                   // exhausted: drop it from the queue (its owner may have done so already)
                lock_guard<mutex> l(m);
                if (!q.empty() && q.front() == t) q.pop_front();
                continue;
            }
            run(*t, i);
            if (t->done.fetch_add(1) + 1 == t->n) { lock_guard<mutex> l(m); cvDone.notify_all(); }
        }
    }
    vector<thread> workers; mutex m; condition_variable cv, cvDone; deque<shared_ptr<Task>> q; bool stop = false;
};

// What a device needs to score the chunks of ONE manifest group (its own background -> its own P and thresholds,
// pwmscan.cpp:594-616).  Jobs carry a pointer to it, so the reader can already parse group g+1 while the GPUs and the
// formatting threads still work on group g: the run is one pipeline, not one per group.
struct GroupParams {
    const Species* species = nullptr;
    uint64_t id = 0;
    vector<float> P; int ldp = 0; vector<int32_t> len; vector<float> thr;
    size_t maxNameLen = 0;                         // longest "<sequence>\tblamm\t<motif>" prefix, for sizing the text buffers
};

struct Job {
    unique_ptr<char[]> chars;                      // not a vector: no zero fill before the parallel copy
    // packed hand-over (default): 2 bits per character + the zero mask, produced by the parser threads (FastaStream::packChunk);
    // `chars` stays empty.  BLAMM_B200_ASCII=1 sends the characters instead and lets the device pack them.
    unique_ptr<uint32_t[]> codes, zmask;
    bool hasZero = false;
    vector<uint64_t> fragStarts;
    vector<Fragment> frags;
    uint64_t nTotal = 0, nPayload = 0;
    uint64_t seq = 0;                              // index of the chunk in the run's stream: the file holds the chunks in this order
    shared_ptr<const GroupParams> group;
};

// A column's share of an occurrence line (pwmscan.cpp:88-95):  <sequence> mid <pos> \t <pos + size> \t <score> tail
struct MotifText {
    string name; uint64_t size = 0; bool revComp = false;
    string mid;                                    // "\tblamm\t<motif>\t"
    char tail[8];                                  // "\t+\t.\t.\n" or "\t-\t.\t.\n" (7 characters) + one spare byte: one 8-byte store
    MotifText(const string& n, uint64_t sz, bool rc) : name(n), size(sz), revComp(rc), mid("\tblamm\t" + n + "\t")
    {
        memcpy(tail, rc ? "\t-\t.\t.\n" : "\t+\t.\t.\n", 8);             // (the 8th byte is the literal's terminator; the next line overwrites it)
    }
};
// One line at p (the caller's buffer has room for the worst case, formatRange / formatBuckets); returns the end of the line.
inline char* formatLine(char* p, const string& seqName, const MotifText& m, uint64_t seqPos, float score)
{
    memcpy(p, seqName.data(), seqName.size()); p += seqName.size();
    memcpy(p, m.mid.data(), m.mid.size()); p += m.mid.size();
    p = to_chars(p, p + 24, (unsigned long long)seqPos).ptr; *p++ = '\t';
    p = to_chars(p, p + 24, (unsigned long long)(seqPos + m.size)).ptr; *p++ = '\t';
    p += formatScore(p, score);
    memcpy(p, m.tail, 8);
    return p + 7;
}
// Buffers for formatted text are recycled: a fresh multi-megabyte allocation is mapped and unmapped by malloc every time, and every
// page of it faults (and is zeroed by the kernel) again -- as much memory traffic as the formatting itself.
class TextPool {
public:
    char* get(size_t cap, size_t& got)
    {
        {
            lock_guard<mutex> l(m);
            auto it = pool.lower_bound(cap);
            if (it != pool.end() && it->first <= 2 * cap + (1u << 20)) { char* p = it->second; got = it->first; held -= got; pool.erase(it); return p; }
        }
        got = cap;
        return new char[cap];
    }
    void put(char* p, size_t cap)
    {
        {
            lock_guard<mutex> l(m);
            if (held + cap <= limit) { pool.emplace(cap, p); held += cap; return; }
        }
        delete[] p;
    }
    ~TextPool() { for (auto& kv : pool) delete[] kv.second; }
private:
    mutex m; multimap<size_t, char*> pool; size_t held = 0; const size_t limit = 48ull << 30;
} gTextPool;
// A piece of formatted occurrence text.  Not a std::string: resize() would zero-fill the worst-case capacity (three times the final
// size) before the lines are written -- 113 GB of memset for configs[2]'s 756 M occurrences.
struct Text {
    char* p = nullptr; size_t cap = 0, n = 0;
    Text() = default;
    Text(const Text&) = delete; Text& operator=(const Text&) = delete;
    Text(Text&& o) noexcept : p(o.p), cap(o.cap), n(o.n) { o.p = nullptr; o.cap = o.n = 0; }
    Text& operator=(Text&& o) noexcept { if (this != &o) { release(); p = o.p; cap = o.cap; n = o.n; o.p = nullptr; o.cap = o.n = 0; } return *this; }
    ~Text() { release(); }
    void release() { if (p) gTextPool.put(p, cap); p = nullptr; cap = n = 0; }
    char* alloc(size_t want) { release(); p = gTextPool.get(want, cap); n = 0; return p; }
    const char* data() const { return p; }
    size_t size() const { return n; }
};

struct ScanShared {
    vector<MotifText> mtext;
    int fd = -1;                                   // occurrence file (byte ranges handed out in chunk order)
    // Write through shared mappings where the file lives in tmpfs, with pwrite elsewhere (BLAMM_B200_WRITER=mmap|pwrite forces one):
    // measured on a 16-core B200 box (tools/micro/file_write.cpp, one new file, GB/s at 4 / 8 / 16 threads) -- tmpfs: mapping
    // 5.3 / 8.1 / 6.7, pwrite 3.8 / 3.8 / 3.8 (the inode lock serialises buffered writes); ext4: mapping 4.2, pwrite 5.1 at 16.
    // More than eight threads faulting pages of one file get in each other's way, so a chunk is copied on at most eight (copyToFile).
    atomic<bool> useMmap{true};
    void chooseWriter()
    {
        const char* e = getenv("BLAMM_B200_WRITER");
        if (e && string(e) == "pwrite") { useMmap = false; return; }
        if (e && string(e) == "mmap") { useMmap = true; return; }
        struct statfs st;
        useMmap = fd >= 0 && fstatfs(fd, &st) == 0 && (unsigned long)st.f_type == 0x01021994ul;      // TMPFS_MAGIC
        installBusHandler();
    }
    // A store into the mapping of a tmpfs file that cannot get its page (the file system is full) raises SIGBUS: end the run
    // with the message and the exit code a failed write() gives, not with a core dump.
    static void installBusHandler()
    {
        static std::once_flag once;
        std::call_once(once, [] {
            struct sigaction sa, old;
            if (sigaction(SIGBUS, nullptr, &old) != 0 || old.sa_handler != SIG_DFL) return;      // somebody else's handler stays
            memset(&sa, 0, sizeof sa);
            sa.sa_handler = [](int) {
                static const char msg[] = "Cannot write to the occurrence file (no space left on the device?), or a mapped input file shrank during the run\n";
                if (write(2, msg, sizeof msg - 1) < 0) {}
                _exit(EXIT_FAILURE);
            };
            sigaction(SIGBUS, &sa, nullptr);
        });
    }
    size_t copyThreads = getenv("BLAMM_B200_COPY_THREADS") ? (size_t)max(1, atoi(getenv("BLAMM_B200_COPY_THREADS"))) : 0;      // 0: 8 (mapping) / 16 (pwrite)
    WorkPool* pool = nullptr;
    // The emitter: ONE thread puts the chunks' text into the file, chunk after chunk in the order the byte ranges were handed
    // out, each chunk on the emitter's OWN copy threads -- so the device workers go on formatting the next chunk while the
    // previous one is being written, no formatting thread ever waits for a turn to copy, and no copy waits behind formatting
    // work in the shared pool's queue.  At most two chunks wait behind the one being written (their text is memory).
    struct EmitJob { vector<Text> text; uint64_t at = 0, bytes = 0; };
    mutex eMutex; condition_variable eCv; deque<EmitJob> eQueue; bool eStop = false, eBusy = false; thread emitter;
    unique_ptr<WorkPool> copyPool;
    void enqueueEmit(EmitJob&& j);             // cli.cpp: below emitText
    void drainEmitter();                       // waits until everything queued is in the file, then ends the thread
    ~ScanShared() { drainEmitter(); }
    atomic<uint64_t> totMatches{0};
    // job queue (producer = FASTA reader, consumers = one thread per GPU)
    mutex qMutex; condition_variable qCv;
    deque<unique_ptr<Job>> queue; bool done = false; size_t maxQueue = 2;
    // Output: the chunks land in the file in STREAM order whatever GPU scored them (Job::seq) -- the per-GPU hit lists are merged
    // in the order of the input, so the file does not depend on -g.  There is no reorder buffer: when a chunk's text is ready
    // its owner waits for its turn (all earlier chunks have taken their place), takes the next `bytes` of the file and leaves
    // the text to the emitter, which writes the pieces in parallel (the reference formats outside and writes inside its output
    // mutex, pwmscan.cpp:88-101).
    mutex oMutex; condition_variable oCv;
    uint64_t nextOut = 0, fileOffset = 0;
    string error; atomic<bool> failed{false};
    void fail(const string& what)
    {
        { lock_guard<mutex> l(qMutex); if (!failed.exchange(true)) error = what; }
        qCv.notify_all();
        { lock_guard<mutex> l(oMutex); }
        oCv.notify_all();
        { lock_guard<mutex> l(eMutex); }
        eCv.notify_all();
    }
};

// hits of one block -> text lines (reference format: pwmscan.cpp:88-95), in (position, column) order.
// The block is cut into position ranges; each formatting thread sorts and formats its range, the pieces go to the file in order.
// H = b200scan_hit12 under BLAMM_B200_HITS=12 (the round-1 hand-over: unordered records, host sort); the default hand-over is the
// ordered 8-byte records of formatBuckets below.
// (position, column) order by LSD radix sort: the hit list of a block comes back in no particular order, and a comparison sort
// was half of the formatting time.  Digits of <= 12 bits over the column and then over the position relative to the smallest
// one in the list (3 passes for 1800 columns and a 6 M position range); counting passes are stable.
template <class H>
void sortHits(std::vector<H>& hits)
{
    const size_t n = hits.size();
    auto less = [](const H& a, const H& b) { return a.pos != b.pos ? a.pos < b.pos : a.col < b.col; };
    if (n < 4096) { sort(hits.begin(), hits.end(), less); return; }
    uint64_t pLo = hits[0].pos, pHi = hits[0].pos; uint32_t cHi = 0;
    for (const auto& h : hits) { pLo = min<uint64_t>(pLo, h.pos); pHi = max<uint64_t>(pHi, h.pos); cHi = max(cHi, h.col); }
    auto bitsOf = [](uint64_t v) { unsigned b = 0; while (v) { b++; v >>= 1; } return b; };
    std::vector<H> tmp(n);
    H* src = hits.data(); H* dst = tmp.data();
    std::vector<size_t> count;
    auto passes = [&](unsigned bits, auto key) {
        if (!bits) return;
        const unsigned np = (bits + 11) / 12, width = (bits + np - 1) / np;
        for (unsigned pass = 0; pass < np; pass++) {
            const unsigned shift = pass * width; const uint64_t mask = (1ull << width) - 1;
            count.assign((size_t)mask + 2, 0);
            for (size_t i = 0; i < n; i++) count[((key(src[i]) >> shift) & mask) + 1]++;
            for (size_t b = 1; b <= mask; b++) count[b] += count[b - 1];
            for (size_t i = 0; i < n; i++) dst[count[(key(src[i]) >> shift) & mask]++] = src[i];
            std::swap(src, dst);
        }
    };
    passes(bitsOf(cHi), [](const H& h) { return (uint64_t)h.col; });
    passes(bitsOf(pHi - pLo), [pLo](const H& h) { return (uint64_t)h.pos - pLo; });
    if (src != hits.data()) hits.swap(tmp);
}

template <class H>
void formatRange(const ScanShared& sh, const Job& job, std::vector<H>& hits, Text& text)
{
    sortHits(hits);
    // worst-case line: names + 2 positions of <= 20 digits + score (<= 16) + 5 tabs + strand + "\t.\t.\n"
    char* const base = text.alloc(hits.size() * (job.group->maxNameLen + 96) + 64);
    char* p = base;
    const vector<string>& names = job.group->species->seqNames;
    size_t f = 0;
    const string* sn = nullptr;                   // name of fragment f's record, looked up when the fragment changes
    for (const auto& h : hits) {
        if (!sn || (f + 1 < job.frags.size() && job.frags[f + 1].streamPos <= h.pos)) {
            while (f + 1 < job.frags.size() && job.frags[f + 1].streamPos <= h.pos) f++;
            sn = &names.at(job.frags[f].seqIdx);                                     // (.at: a record the .dict does not know ends the run with an error)
        }
        const Fragment& fr = job.frags[f];
        p = formatLine(p, *sn, sh.mtext[h.col], fr.seqPos + (h.pos - fr.streamPos), h.score);
    }
    text.n = (size_t)(p - base);
}

// The text of one chunk goes to its byte range of the file: pieces of at most 16 MiB, through a shared mapping of the range on
// eight threads where the file lives in tmpfs (buffered write() calls on one file serialise on the inode lock -- measured: 16
// threads of pwrite = one thread's 3.7 GB/s -- page faults of a mapping do not, but more than eight threads faulting pages of
// one file get in each other's way), with pwrite on sixteen elsewhere (BLAMM_B200_COPY_THREADS).  Runs on the emitter thread.
void copyToFile(ScanShared& sh, ScanShared::EmitJob& j)
{
    const double t0 = now();
    bool mapped = sh.useMmap;
    struct Piece { const char* p; size_t n; uint64_t at; };
    vector<Piece> pieces;
    {
        uint64_t o = j.at;
        for (const auto& t : j.text) {
            for (size_t q = 0; q < t.size(); q += (16u << 20)) pieces.push_back({t.data() + q, min<size_t>(16u << 20, t.size() - q), o + q});
            o += t.size();
        }
    }
    atomic<bool> bad{false};
    char* base = nullptr; uint64_t mapAt = 0; size_t mapLen = 0;
    if (mapped) {
        const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
        mapAt = j.at / page * page; mapLen = (size_t)(j.at + j.bytes - mapAt);
        void* m = mmap(nullptr, mapLen, PROT_READ | PROT_WRITE, MAP_SHARED, sh.fd, (off_t)mapAt);
        if (m == MAP_FAILED) { mapped = false; sh.useMmap = false; } else base = static_cast<char*>(m);
    }
    // strand i takes pieces i, i + S, ... on the emitter's own threads
    const size_t S = min(pieces.size(), sh.copyPool->size());
    sh.copyPool->parallel(S, [&](size_t strand) {
        for (size_t i = strand; i < pieces.size(); i += S) {
            if (mapped) { memcpy(base + (pieces[i].at - mapAt), pieces[i].p, pieces[i].n); continue; }
            const char* p = pieces[i].p; size_t left = pieces[i].n; uint64_t o = pieces[i].at;
            while (left) {
                const ssize_t w = pwrite(sh.fd, p, left, (off_t)o);
                if (w <= 0) { if (w < 0 && errno == EINTR) continue; bad = true; return; }
                p += w; left -= (size_t)w; o += (uint64_t)w;
            }
        }
    });
    if (base) munmap(base, mapLen);
    if (bad) sh.fail("Cannot write to the occurrence file");
    gTimer.add("file write (emitter thread, overlaps the formatting)", now() - t0);
}

void ScanShared::enqueueEmit(EmitJob&& j)
{
    unique_lock<mutex> l(eMutex);
    if (!emitter.joinable()) {
        eStop = false;
        copyPool.reset(new WorkPool(copyThreads ? copyThreads : (useMmap ? 8 : 16)));
        emitter = thread([this] {
            for (;;) {
                EmitJob job;
                {
                    unique_lock<mutex> q(eMutex);
                    eCv.wait(q, [&] { return eStop || !eQueue.empty(); });
                    if (eQueue.empty()) return;
                    job = std::move(eQueue.front()); eQueue.pop_front();
                    eBusy = true;
                }
                eCv.notify_all();
                try { if (!failed) copyToFile(*this, job); } catch (const exception& e) { fail(e.what()); }
                job.text.clear();                                  // (the buffers go back to the pool before the next wait)
                { lock_guard<mutex> q(eMutex); eBusy = false; }
                eCv.notify_all();
            }
        });
    }
    eCv.wait(l, [&] { return eQueue.size() < 2 || failed; });
    if (failed) return;
    eQueue.push_back(std::move(j));
    l.unlock();
    eCv.notify_all();
}

void ScanShared::drainEmitter()
{
    {
        unique_lock<mutex> l(eMutex);
        if (!emitter.joinable()) return;
        eCv.wait(l, [&] { return (eQueue.empty() && !eBusy) || failed; });
        eStop = true;
    }
    eCv.notify_all();
    emitter.join();
}

// formatted text of one chunk -> its place in the file: wait for the chunk's turn (all earlier chunks have taken their byte ranges),
// take the next `bytes` of the file, and leave the copying to the emitter
void emitText(ScanShared& sh, const Job& job, vector<Text>& text, uint64_t n)
{
    double t0 = now();
    ScanShared::EmitJob j;
    for (const auto& t : text) j.bytes += t.size();
    {
        unique_lock<mutex> lock(sh.oMutex);
        sh.oCv.wait(lock, [&] { return job.seq == sh.nextOut || sh.failed; });
        if (sh.failed) return;
        j.at = sh.fileOffset;
        sh.fileOffset += j.bytes;
        // (the file grows here, in chunk order, so that every chunk can map its own byte range)
        if (sh.useMmap && j.bytes && ftruncate(sh.fd, (off_t)sh.fileOffset) != 0) sh.useMmap = false;
        sh.nextOut++;
    }
    sh.oCv.notify_all();
    sh.totMatches += n;
    gTimer.add("wait for the chunk's turn in the file (wall)", now() - t0); t0 = now();
    if (!j.bytes) return;
    j.text = std::move(text);
    text.clear();
    sh.enqueueEmit(std::move(j));
    gTimer.add("wait for room behind the emitter (wall)", now() - t0);
}

template <class H>
void writeHits(ScanShared& sh, const Job& job, const H* hits, uint64_t n, size_t threads)
{
    double t0 = now();
    const size_t T = std::max<size_t>(1, std::min<size_t>(threads, n / 50000 + 1));
    vector<vector<H>> part(T);
    if (T == 1) part[0].assign(hits, hits + n);
    else {
        // position ranges of equal width; counting and scattering are themselves split over the threads (hit order from
        // the device is arbitrary, so every thread scans its slice of the list and appends to per-(thread, range) bins)
        const uint64_t span = job.nPayload / T + 1;
        vector<vector<vector<H>>> bins(T, vector<vector<H>>(T));
        sh.pool->parallel(T, [&](size_t t) {
            const uint64_t lo = n * t / T, hi = n * (t + 1) / T;
            for (auto& b : bins[t]) b.reserve((hi - lo) / T + (hi - lo) / (4 * T) + 16);
            for (uint64_t i = lo; i < hi; i++) bins[t][hits[i].pos / span].push_back(hits[i]);
        });
        sh.pool->parallel(T, [&](size_t r) {
            size_t c = 0;
            for (size_t t = 0; t < T; t++) c += bins[t][r].size();
            part[r].reserve(c);
            for (size_t t = 0; t < T; t++) part[r].insert(part[r].end(), bins[t][r].begin(), bins[t][r].end());
        });
    }
    gTimer.add("partition hits (wall)", now() - t0); t0 = now();
    vector<Text> text(T);
    sh.pool->parallel(T, [&](size_t t) { formatRange<H>(sh, job, part[t], text[t]); });
    gTimer.add("sort + format (wall)", now() - t0);
    emitText(sh, job, text, n);
}

// Ordered 8-byte records (B200SCAN_HITS_8): the device has already put the block's hits in (position, column) order and
// grouped them in buckets of 256 positions, so the host neither partitions nor sorts: every formatting thread takes a run of
// buckets holding about n / T hits.
void formatBuckets(const ScanShared& sh, const Job& job, const b200scan_hit8* hits, const uint32_t* bucketStart, uint64_t b0, uint64_t b1,
                   Text& text, uint64_t posOffset = 0)
{
    const uint64_t n = bucketStart[b1] - bucketStart[b0];
    char* const base = text.alloc(n * (job.group->maxNameLen + 96) + 64);
    char* p = base;
    const vector<string>& names = job.group->species->seqNames;
    size_t f = 0;
    const string* sn = nullptr;                   // name of fragment f's record, looked up when the fragment changes
    for (uint64_t b = b0; b < b1; b++) {
        const uint64_t bucketPos = posOffset + (b << B200SCAN_BUCKET_SHIFT);
        for (uint32_t i = bucketStart[b]; i < bucketStart[b + 1]; i++) {
            const uint64_t pos = bucketPos + (hits[i].key >> 24);
            if (!sn || (f + 1 < job.frags.size() && job.frags[f + 1].streamPos <= pos)) {
                while (f + 1 < job.frags.size() && job.frags[f + 1].streamPos <= pos) f++;
                sn = &names.at(job.frags[f].seqIdx);                                 // (.at: a record the .dict does not know ends the run with an error)
            }
            const Fragment& fr = job.frags[f];
            p = formatLine(p, *sn, sh.mtext[hits[i].key & 0xFFFFFFu], fr.seqPos + (pos - fr.streamPos), hits[i].score);
        }
    }
    text.n = (size_t)(p - base);
}

// format the ordered hit list of (a piece of) a chunk: appends the text pieces to `text`
void formatHits8(ScanShared& sh, const Job& job, const b200scan_hit8* hits, uint64_t n, const uint32_t* bucketStart, uint64_t nBuckets, size_t threads,
                 vector<Text>& text, uint64_t posOffset = 0)
{
    const double t0 = now();
    // a few pieces per thread: the pool is shared by all the GPUs' chunks, short pieces even out the load
    const size_t T = std::max<size_t>(1, std::min<size_t>(2 * threads, n / 50000 + 1));
    vector<uint64_t> cut(T + 1, nBuckets);
    cut[0] = 0;
    for (size_t t = 1; t < T; t++)                   // first bucket whose start reaches t/T of the hits
        cut[t] = (uint64_t)(std::lower_bound(bucketStart, bucketStart + nBuckets, (uint32_t)(n * t / T)) - bucketStart);
    const size_t first = text.size();
    text.resize(first + T);
    sh.pool->parallel(T, [&](size_t t) { formatBuckets(sh, job, hits, bucketStart, cut[t], cut[t + 1], text[first + t], posOffset); });
    gTimer.add("format ordered hits (wall)", now() - t0);
}

void writeHits8(ScanShared& sh, const Job& job, const b200scan_hit8* hits, uint64_t n, const uint32_t* bucketStart, uint64_t nBuckets, size_t threads)
{
    vector<Text> text;
    formatHits8(sh, job, hits, n, bucketStart, nBuckets, threads, text);
    emitText(sh, job, text, n);
}

// A chunk whose hits do not fit the device buffers (b200scan_collect8 -> B200SCAN_ENOMEM: thresholds so low that a large part of
// all windows are occurrences): the payload range [lo, hi) is scored again in two halves, recursively, each half a block of its
// own with the halo behind it; the text comes out in position order.  The reference streams any hit density to disk block by
// block (250,000 characters, pwmscan.cpp:258); here only such dense chunks fall back to smaller blocks.
bool scanSplit(ScanShared& sh, b200scan_ctx* ctx, const Job& job, uint64_t lo, uint64_t hi, uint64_t halo, bool foldLower, size_t threads,
               vector<Text>& text, uint64_t& nHits, string& err, size_t devIndex)
{
    const uint64_t nTotal = std::min<uint64_t>(job.nTotal - lo, hi - lo + halo);
    vector<uint64_t> frag;
    for (uint64_t f : job.fragStarts) if (f > lo && f < lo + nTotal) frag.push_back(f - lo);
    const int rc = job.codes
        ? b200scan_submit_packed(ctx, 0, job.codes.get() + lo / 16, job.hasZero ? job.zmask.get() + lo / 32 : nullptr, nTotal, hi - lo, frag.data(), frag.size())
        : b200scan_submit_ascii(ctx, 0, job.chars.get() + lo, nTotal, hi - lo, frag.data(), frag.size(), foldLower ? B200SCAN_LOWER_FOLD : B200SCAN_LOWER_ZERO);
    if (rc != B200SCAN_OK) { err = b200scan_last_error(ctx); return false; }
    const b200scan_hit8* hits = nullptr; const uint32_t* bucketStart = nullptr; uint64_t n = 0, nb = 0;
    b200scan_timing tm{};
    const int rc2 = b200scan_collect8(ctx, 0, &hits, &n, &bucketStart, &nb, &tm);
    if (rc2 == B200SCAN_OK) {
        gStats.add(devIndex, tm, hi - lo);
        formatHits8(sh, job, hits, n, bucketStart, nb, threads, text, lo);
        nHits += n;
        return true;
    }
    if (rc2 != B200SCAN_ENOMEM || hi - lo < 2048) { err = b200scan_last_error(ctx); return false; }
    const uint64_t mid = lo + ((hi - lo) / 2 + 31) / 32 * 32;           // code and mask words start at multiples of 32 characters
    return scanSplit(sh, ctx, job, lo, mid, halo, foldLower, threads, text, nHits, err, devIndex) &&
           scanSplit(sh, ctx, job, mid, hi, halo, foldLower, threads, text, nHits, err, devIndex);
}

// One thread per GPU for the whole run.  The context (CUDA initialisation ~0.5-1.5 s) is created on its own thread while the host
// still loads the inputs (`ready` carries the error text of a failed creation); a job of another manifest group than the one loaded
// makes the worker drain its chunks in flight and load that group's P and thresholds (b200scan_set_motifs).
void deviceWorker(ScanShared& sh, int engine, bool foldLower, b200scan_ctx** ctxSlot, shared_future<string> ready, size_t devIndex, size_t threads,
                  uint64_t halo)
{
    b200scan_ctx*& ctx = *ctxSlot;
    auto die = [&](const string& what) { sh.fail(what); };
    {   // the context is being created since the start of the run (runScan: CtxPool), normally it is ready by now
        const double t0 = now();
        const string err = ready.get();
        gTimer.add("wait for b200scan_create", now() - t0);
        if (!err.empty() || !ctx) { die(err.empty() ? "CUDA error: no context" : err); return; }
    }
    const char* hitEnv = getenv("BLAMM_B200_HITS");
    const int hitFormat = (hitEnv && atoi(hitEnv) == 12) || sh.mtext.size() > (1u << 24) ? B200SCAN_HITS_12 : B200SCAN_HITS_8;
    if (b200scan_set_engine(ctx, engine) != B200SCAN_OK || b200scan_set_hit_format(ctx, hitFormat) != B200SCAN_OK) {
        die(string("CUDA error: ") + b200scan_last_error(ctx)); return;
    }
    shared_ptr<const GroupParams> loaded;
    // Three slots: while the hits of chunk k-1 come down and are formatted, chunk k is being scored and chunk k+1 goes up.
    // BLAMM_B200_HITS=12 keeps the unordered 12-byte records and the host sort (diagnostic / comparison).
    unique_ptr<Job> inFlight[B200SCAN_NUM_SLOTS];
    int head = 0, tail = 0, nFlight = 0;               // slots tail .. head-1 (mod 3) hold chunks, oldest first
    // dense chunks (B200SCAN_ENOMEM at collect) waiting to be scored in halves, and the chunks collected behind them
    struct Held { unique_ptr<Job> job; vector<b200scan_hit8> hits; vector<uint32_t> buckets; };
    vector<unique_ptr<Job>> pendingSplit; vector<Held> held;
    auto collectOldest = [&]() -> bool {
        const int s = tail;
        const double tc = now();
        b200scan_timing tm{};
        if (hitFormat == B200SCAN_HITS_8) {
            const b200scan_hit8* hits = nullptr; const uint32_t* bucketStart = nullptr; uint64_t n = 0, nb = 0;
            const int rc = b200scan_collect8(ctx, s, &hits, &n, &bucketStart, &nb, &tm);
            if (rc == B200SCAN_ENOMEM) {
                // too dense for the device buffers: finish the younger chunks in flight (they may be as dense: same treatment, in
                // order), then score this one in halves.  All slots are free while a chunk is split.
                pendingSplit.push_back(std::move(inFlight[s]));
                tail = (tail + 1) % B200SCAN_NUM_SLOTS; nFlight--;
                return !sh.failed;
            }
            if (rc != B200SCAN_OK) { die(string("CUDA error: ") + b200scan_last_error(ctx)); return false; }
            gTimer.add("b200scan_collect (wait GPU)", now() - tc);
            if (!pendingSplit.empty()) {                      // keep the stream order: a younger chunk waits behind the split ones (rare path: copy its hits)
                Held h; h.job = std::move(inFlight[s]); h.hits.assign(hits, hits + n); h.buckets.assign(bucketStart, bucketStart + nb + 1);
                held.push_back(std::move(h));
                gStats.add(devIndex, tm, held.back().job->nPayload);
                tail = (tail + 1) % B200SCAN_NUM_SLOTS; nFlight--;
                return !sh.failed;
            }
            writeHits8(sh, *inFlight[s], hits, n, bucketStart, nb, threads);
        } else {
            const b200scan_hit12* hits = nullptr; uint64_t n = 0;
            if (b200scan_collect12(ctx, s, &hits, &n, &tm) != B200SCAN_OK) { die(string("CUDA error: ") + b200scan_last_error(ctx)); return false; }
            gTimer.add("b200scan_collect (wait GPU)", now() - tc);
            writeHits(sh, *inFlight[s], hits, n, threads);
        }
        gStats.add(devIndex, tm, inFlight[s]->nPayload);
        inFlight[s].reset();
        tail = (tail + 1) % B200SCAN_NUM_SLOTS; nFlight--;
        return !sh.failed;
    };
    // dense chunks: drain what is in flight, score them in halves (slot 0, synchronously), then release the chunks held behind them.
    // Everything this worker holds is emitted in increasing chunk order, so nobody waits for a chunk that cannot come.
    auto resolveSplits = [&]() -> bool {
        while (nFlight > 0) if (!collectOldest()) return false;
        struct Item { uint64_t seq; int kind; size_t idx; };
        vector<Item> order;
        for (size_t i = 0; i < pendingSplit.size(); i++) order.push_back({pendingSplit[i]->seq, 0, i});
        for (size_t i = 0; i < held.size(); i++) order.push_back({held[i].job->seq, 1, i});
        sort(order.begin(), order.end(), [](const Item& a, const Item& b) { return a.seq < b.seq; });
        for (const auto& it : order) {
            if (it.kind == 0) {
                const Job& job = *pendingSplit[it.idx];
                vector<Text> text; uint64_t n = 0; string err;
                const uint64_t mid = (job.nPayload / 2 + 31) / 32 * 32;
                const bool ok = job.nPayload < 4096 ? false
                              : scanSplit(sh, ctx, job, 0, mid, halo, foldLower, threads, text, n, err, devIndex) &&
                                scanSplit(sh, ctx, job, mid, job.nPayload, halo, foldLower, threads, text, n, err, devIndex);
                if (!ok) { die("CUDA error: " + (err.empty() ? string("chunk too dense for the device buffers") : err)); return false; }
                emitText(sh, job, text, n);
            } else {
                const Held& h = held[it.idx];
                writeHits8(sh, *h.job, h.hits.data(), h.hits.size(), h.buckets.data(), h.buckets.size() - 1, threads);
            }
            if (sh.failed) return false;
        }
        pendingSplit.clear(); held.clear();
        return true;
    };
    try {
    while (!sh.failed) {
        if (!pendingSplit.empty() && !resolveSplits()) break;
        unique_ptr<Job> job;
        bool drainOne = false;
        {
            unique_lock<mutex> l(sh.qMutex);
            // with chunks in flight and nothing to submit, use the time to collect the oldest one instead of waiting
            if (sh.queue.empty() && !sh.done && !sh.failed && nFlight > 0) drainOne = true;
            else {
                sh.qCv.wait(l, [&] { return !sh.queue.empty() || sh.done || sh.failed; });
                if (sh.failed) break;
                if (sh.queue.empty()) break;                 // done
                job = std::move(sh.queue.front()); sh.queue.pop_front();
                sh.qCv.notify_all();
            }
        }
        if (drainOne) { if (!collectOldest()) break; continue; }
        if (job->group != loaded) {                          // next manifest group: its own P and thresholds
            bool ok = true;
            while (nFlight > 0 && ok) ok = collectOldest();
            if (ok && !pendingSplit.empty()) ok = resolveSplits();       // (dense chunks of the old group are re-scored with ITS motifs)
            if (!ok) break;
            const GroupParams& g = *job->group;
            if (b200scan_set_motifs(ctx, g.P.data(), g.ldp, (int32_t)g.len.size(), g.len.data(), g.thr.data()) != B200SCAN_OK) {
                die(string("CUDA error: ") + b200scan_last_error(ctx)); break;
            }
            loaded = job->group;
        }
        if (nFlight == B200SCAN_NUM_SLOTS && !collectOldest()) break;
        const int rc = job->codes
            ? b200scan_submit_packed(ctx, head, job->codes.get(), job->hasZero ? job->zmask.get() : nullptr, job->nTotal, job->nPayload,
                                     job->fragStarts.data(), job->fragStarts.size())
            : b200scan_submit_ascii(ctx, head, job->chars.get(), job->nTotal, job->nPayload, job->fragStarts.data(),
                                    job->fragStarts.size(), foldLower ? B200SCAN_LOWER_FOLD : B200SCAN_LOWER_ZERO);
        if (rc != B200SCAN_OK) {
            die(string("CUDA error: ") + b200scan_last_error(ctx)); break;
        }
        inFlight[head] = std::move(job);
        head = (head + 1) % B200SCAN_NUM_SLOTS; nFlight++;
        if (nFlight == B200SCAN_NUM_SLOTS && !collectOldest()) break;     // overlap: format chunk k-2 while k-1 is scored and k goes up
    }
    while (nFlight > 0 && !sh.failed) if (!collectOldest()) break;
    if (!pendingSplit.empty() && !sh.failed) resolveSplits();
    } catch (const exception& e) {                       // e.g. a record the .dict does not know (formatting runs on the pool: WorkPool rethrows here)
        die(e.what());
    }
}

// `blamm-b200 selftest-writer [hits] [threads]` (no GPU): the occurrence writer -- partition, radix sort, formatting, writer
// thread -- on a synthetic hit list of both record types, against a plain restatement (std::sort, upper_bound, snprintf "%g").
template <class H>
bool writerSelfTest(size_t nHits, size_t threads, const char* label)
{
    std::mt19937_64 rng(12345 + nHits);
    MotifSet ms;
    const size_t nCols = 37;
    for (size_t c = 0; c < nCols; c++) {
        Motif m; m.name = "MA" + to_string(1000 + c / 2) + "." + to_string(1 + c % 3); m.pfm.resize(5 + c % 17); m.revComp = c & 1;
        ms.motifs.push_back(m);
    }
    Species sp; sp.name = "syn";
    sp.seqNames = {"chr1", "chr2_with_a_longer_name", "s3"};
    auto group = make_shared<GroupParams>();
    group->species = &sp; group->maxNameLen = 23 + 8;
    WorkPool workPool(threads);
    auto openOut = [](const string& path) { return open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644); };
    auto fillShared = [&](ScanShared& sh, int fd) {
        for (const auto& m : ms.motifs) sh.mtext.emplace_back(m.name, (uint64_t)m.size(), m.revComp);
        sh.fd = fd; sh.pool = &workPool;
        sh.chooseWriter();
    };
    Job job;
    job.group = group;
    job.nPayload = 3000000; job.nTotal = job.nPayload + 40;
    job.frags = {{0, 0, 1234567890123ull}, {700001, 0, 1234567990123ull}, {1500000, 1, 0}, {1500007, 1, 19}, {2999990, 2, 5}};
    // distinct (position, column) pairs, shuffled
    const uint64_t universe = job.nPayload * nCols, g = max<uint64_t>(1, universe / max<size_t>(1, nHits));
    vector<H> hits;
    for (size_t i = 0; i < nHits && i * g < universe; i++) {
        const uint64_t key = i * g + rng() % g;
        H h; h.pos = (decltype(h.pos))(key / nCols); h.col = (uint32_t)(key % nCols);
        const int kind = (int)(rng() % 8);
        const double u = (double)(rng() % 2000001) / 1e6 - 1.0;                       // [-1, 1]
        h.score = kind == 0 ? (float)(u * 1e-3) : kind == 1 ? (float)(u * 2e6) : kind == 2 ? (float)((int)(u * 40)) : (float)(u * 30.0);
        hits.push_back(h);
    }
    shuffle(hits.begin(), hits.end(), rng);
    // restatement
    vector<H> inOrder(hits);
    sort(inOrder.begin(), inOrder.end(), [](const H& a, const H& b) { return a.pos != b.pos ? a.pos < b.pos : a.col < b.col; });
    string want;
    char line[512];
    for (const auto& h : inOrder) {
        auto it = upper_bound(job.frags.begin(), job.frags.end(), (uint64_t)h.pos, [](uint64_t p, const Fragment& f) { return p < f.streamPos; }) - 1;
        const unsigned long long seqPos = it->seqPos + (h.pos - it->streamPos);
        const Motif& m = ms.motifs[h.col];
        const int n = snprintf(line, sizeof line, "%s\tblamm\t%s\t%llu\t%llu\t%g\t%c\t.\t.\n", sp.seqNames[it->seqIdx].c_str(), m.name.c_str(),
                               seqPos, seqPos + (unsigned long long)m.size(), (double)h.score, m.revComp ? '-' : '+');
        want.append(line, (size_t)n);
    }
    const string path = string(getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp") + "/blamm_b200_selftest_" + to_string((long)getpid()) + ".txt";
    {
        const int fd = openOut(path);
        ScanShared sh;
        fillShared(sh, fd);
        writeHits(sh, job, hits.data(), hits.size(), threads);
        sh.drainEmitter();
        close(fd);
        if (sh.totMatches != hits.size() || sh.failed) { cerr << label << ": match count differs\n"; return false; }
    }
    ifstream is(path, ios::binary);
    const string got((istreambuf_iterator<char>(is)), istreambuf_iterator<char>());
    remove(path.c_str());
    bool ok = got == want;
    cout << label << ": " << hits.size() << " hits, " << threads << " threads, " << got.size() << " bytes: " << (ok ? "identical" : "DIFFERENT") << "\n";
    if (sizeof(H) == sizeof(b200scan_hit12)) {
        // the same list as the device hands it over under B200SCAN_HITS_8: ordered 8-byte records + bucket index
        const uint64_t nb = (job.nPayload + (1u << B200SCAN_BUCKET_SHIFT) - 1) >> B200SCAN_BUCKET_SHIFT;
        vector<b200scan_hit8> h8(inOrder.size());
        vector<uint32_t> bs(nb + 1, 0);
        for (size_t i = 0; i < inOrder.size(); i++) {
            h8[i].key = (uint32_t)((inOrder[i].pos & 255u) << 24) | inOrder[i].col; h8[i].score = inOrder[i].score;
            bs[(inOrder[i].pos >> B200SCAN_BUCKET_SHIFT) + 1]++;
        }
        for (uint64_t b = 0; b < nb; b++) bs[b + 1] += bs[b];
        {
            const int fd = openOut(path);
            ScanShared sh;
            fillShared(sh, fd);
            writeHits8(sh, job, h8.data(), h8.size(), bs.data(), nb, threads);
            sh.drainEmitter();
            close(fd);
        }
        ifstream is8(path, ios::binary);
        const string got8((istreambuf_iterator<char>(is8)), istreambuf_iterator<char>());
        remove(path.c_str());
        const bool ok8 = got8 == want;
        cout << "b200scan_hit8 : " << h8.size() << " hits, " << threads << " threads, " << got8.size() << " bytes: " << (ok8 ? "identical" : "DIFFERENT") << "\n";
        ok = ok && ok8;
    }
    return ok;
}

// `blamm-b200 selftest-order [chunks] [workers] [threads]` (no GPU): the stream-order merge of `scan -g N`.  `workers` threads play
// the per-GPU workers: chunk k goes to worker k mod workers (or to whoever is free first), every worker formats and emits its
// chunks after a random delay through the product's writeHits8 / emitText, so the chunks become ready in an order that has nothing
// to do with the stream.  The file must hold the chunks in stream order -- byte for byte the text of a single worker doing them
// one after the other -- whatever the number of workers (per-GPU hit lists merged on the host in reference order).
bool orderSelfTest(size_t nChunks, size_t nWorkers, size_t threads)
{
    std::mt19937_64 rng(4711 + nChunks);
    MotifSet ms;
    const size_t nCols = 24;
    for (size_t c = 0; c < nCols; c++) {
        Motif m; m.name = "MB" + to_string(2000 + c / 2) + ".1"; m.pfm.resize(6 + c % 11); m.revComp = c & 1;
        ms.motifs.push_back(m);
    }
    Species sp; sp.name = "syn";
    sp.seqNames = {"chrA", "chrB"};
    auto group = make_shared<GroupParams>();
    group->species = &sp; group->maxNameLen = 4 + 8;
    // the chunks: ordered 8-byte records + bucket index, as b200scan_collect8 returns them; chunk sizes differ (also empty ones)
    struct Chunk { Job job; vector<b200scan_hit8> hits; vector<uint32_t> buckets; uint64_t nb = 0; };
    vector<Chunk> chunks(nChunks);
    uint64_t streamPos = 0;
    for (size_t k = 0; k < nChunks; k++) {
        Chunk& c = chunks[k];
        c.job.group = group; c.job.seq = k;
        c.job.nPayload = 20000 + (rng() % 7) * 9000; c.job.nTotal = c.job.nPayload + 20;
        c.job.frags = {{0, k & 1, streamPos}};
        streamPos += c.job.nPayload;
        c.nb = (c.job.nPayload + (1u << B200SCAN_BUCKET_SHIFT) - 1) >> B200SCAN_BUCKET_SHIFT;
        c.buckets.assign(c.nb + 1, 0);
        const size_t want = (k % 5 == 3) ? 0 : 200 + (size_t)(rng() % 60000);
        const uint64_t universe = c.job.nPayload * nCols, g = max<uint64_t>(1, universe / max<size_t>(1, want));
        for (size_t i = 0; i < want && i * g < universe; i++) {
            const uint64_t key = i * g + rng() % g;                  // ascending (position, column)
            const uint64_t pos = key / nCols; const uint32_t col = (uint32_t)(key % nCols);
            c.hits.push_back({(uint32_t)((pos & 255u) << 24) | col, (float)((double)(rng() % 2000001) / 1e5 - 10.0)});
            c.buckets[(pos >> B200SCAN_BUCKET_SHIFT) + 1]++;
        }
        for (uint64_t b = 0; b < c.nb; b++) c.buckets[b + 1] += c.buckets[b];
    }
    WorkPool workPool(threads);
    const string base = string(getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp") + "/blamm_b200_selftest_order_" + to_string((long)getpid());
    auto run = [&](size_t workers, bool dynamicDeal, const string& path) -> bool {
        const int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
        if (fd < 0) return false;
        ScanShared sh;
        for (const auto& m : ms.motifs) sh.mtext.emplace_back(m.name, (uint64_t)m.size(), m.revComp);
        sh.fd = fd; sh.pool = &workPool;
        sh.chooseWriter();
        atomic<size_t> next{0};
        vector<thread> th;
        for (size_t w = 0; w < workers; w++) th.emplace_back([&, w] {
            std::mt19937_64 r(99 + w);
            // every worker emits ITS chunks in increasing stream order (deviceWorker does: a chunk is collected before the next one
            // submitted behind it); between workers nothing is ordered
            for (size_t k = dynamicDeal ? next.fetch_add(1) : w; k < nChunks; k = dynamicDeal ? next.fetch_add(1) : k + workers) {
                if (workers > 1) std::this_thread::sleep_for(std::chrono::microseconds(r() % 3000));
                const Chunk& c = chunks[k];
                writeHits8(sh, c.job, c.hits.data(), c.hits.size(), c.buckets.data(), c.nb, threads);
            }
        });
        for (auto& t : th) t.join();
        sh.drainEmitter();
        close(fd);
        uint64_t total = 0;
        for (const auto& c : chunks) total += c.hits.size();
        return !sh.failed && sh.totMatches == total && sh.nextOut == nChunks;
    };
    auto slurp = [](const string& path) { ifstream is(path, ios::binary); return string((istreambuf_iterator<char>(is)), istreambuf_iterator<char>()); };
    bool ok = run(1, false, base + "_1.txt");
    const string want = slurp(base + "_1.txt");
    remove((base + "_1.txt").c_str());
    {   // the single-worker file itself: one line per hit, stream positions never decreasing (chunk k's records start where chunk k-1's end)
        uint64_t lines = 0, total = 0, last = 0;
        for (const auto& c : chunks) total += c.hits.size();
        for (size_t at = 0; at < want.size() && ok;) {
            const size_t eol = want.find('\n', at);
            if (eol == string::npos) { ok = false; break; }
            size_t tab = at;
            for (int f = 0; f < 3; f++) tab = want.find('\t', tab) + 1;      // <sequence> blamm <motif> <start> ...
            const uint64_t start = strtoull(want.c_str() + tab, nullptr, 10);
            if (start < last) ok = false;
            last = start; lines++; at = eol + 1;
        }
        ok = ok && lines == total;
        cout << "1 worker: " << lines << " lines of " << total << " hits, " << want.size() << " bytes, positions " << (ok ? "in stream order" : "OUT OF ORDER") << "\n";
    }
    for (int dynamicDeal = 0; dynamicDeal < 2 && ok; dynamicDeal++) {
        const string path = base + "_n.txt";
        const bool ran = run(nWorkers, dynamicDeal != 0, path);
        const string got = slurp(path);
        remove(path.c_str());
        const bool same = ran && got == want;
        cout << nWorkers << " workers, chunks " << (dynamicDeal ? "taken by whoever is free" : "dealt round robin") << ": " << got.size() << " bytes: "
             << (same ? "identical" : "DIFFERENT") << "\n";
        ok = ok && same;
    }
    return ok && !want.empty();
}

int runOrderSelfTest(int argc, char** argv)
{
    const size_t nChunks = argc > 2 ? (size_t)atoll(argv[2]) : 40, nWorkers = argc > 3 ? (size_t)atoll(argv[3]) : 8, threads = argc > 4 ? (size_t)atoll(argv[4]) : 4;
    return orderSelfTest(max<size_t>(1, nChunks), max<size_t>(1, nWorkers), max<size_t>(1, threads)) ? EXIT_SUCCESS : EXIT_FAILURE;
}

int runWriterSelfTest(int argc, char** argv)
{
    const size_t nHits = argc > 2 ? (size_t)atoll(argv[2]) : 300000, threads = argc > 3 ? (size_t)atoll(argv[3]) : 4;
    bool ok = writerSelfTest<b200scan_hit>(nHits, threads, "b200scan_hit  ");
    ok = writerSelfTest<b200scan_hit12>(nHits, threads, "b200scan_hit12") && ok;
    gTimer.report();                                   // BLAMM_B200_TIMING=1: the phases of both runs, summed
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}

} // namespace

void writeStats()
{
    if (gStats.file.empty()) return;
    ofstream js(gStats.file);
    js.precision(9);
    js << "{\"phases_s\": {";
    for (size_t i = 0; i < gTimer.acc.size(); i++) js << (i ? ", " : "") << "\"" << gTimer.acc[i].first << "\": " << gTimer.acc[i].second;
    js << "}, \"devices\": [";
    for (size_t d = 0; d < gStats.dev.size(); d++) {
        const DeviceStats& s = gStats.dev[d];
        js << (d ? ", " : "") << "{\"device\": " << d << ", \"chunks\": " << s.chunks << ", \"characters\": " << s.chars << ", \"hits\": " << s.hits
           << ", \"candidates\": " << s.candidates << ", \"h2d_ms\": " << s.h2d << ", \"pack_ms\": " << s.pack << ", \"score_ms\": " << s.score
           << ", \"rescore_ms\": " << s.rescore << ", \"order_ms\": " << s.order << ", \"d2h_ms\": " << s.d2h << "}";
    }
    js << "], \"columns\": " << gStatsColumns << ", \"matches\": " << gStatsMatches << "}\n";
}

int runScan(int argc, char** argv)
{
    bool foldLower = false, revCompl = false;
    bool absSpec = false, relSpec = false, pSpec = false;
    float absThr = 0.0f, relThr = 0.95f, pvalue = 0.0001f;
    size_t numThreads = thread::hardware_concurrency();
    int gpusWanted = 0, engine = B200SCAN_ENGINE_AUTO;
    string histdir, outputFilename = "occurrences.txt";
    if (argc < 4) { scanUsage(); return EXIT_FAILURE; }
    for (int i = 2; i < argc - 2; i++) {
        string arg(argv[i]);
        const bool hasVal = i + 1 < argc - 2;
        if (arg == "-h" || arg == "--help") { scanUsage(); return EXIT_SUCCESS; }
        else if (arg == "-rc" || arg == "--revcompl") revCompl = true;
        else if (arg == "-s" || arg == "--simple") foldLower = true;
        else if (arg == "-c" || arg == "--cuda") {}
        else if ((arg == "-at" || arg == "--absthreshold") && hasVal) { absSpec = true; absThr = atof(argv[++i]); }
        else if ((arg == "-rt" || arg == "--relthreshold") && hasVal) {
            relSpec = true; relThr = atof(argv[++i]);
            if (relThr < 0.0 || relThr > 1.0) throw runtime_error("The relative threshold should be in range [0..1].");
        } else if ((arg == "-pt" || arg == "--pthreshold") && hasVal) {
            pSpec = true; pvalue = atof(argv[++i]);
            if (pvalue < 0.0 || pvalue > 1.0) throw runtime_error("The p-value should be in range [0..1].");
        } else if ((arg == "-t" || arg == "--numthreads") && hasVal) {
            const int t = atoi(argv[++i]);
            if (t < 1) throw runtime_error("Number of threads must be a non-zero positive number");
            numThreads = t;
        } else if ((arg == "-g" || arg == "--gpus") && hasVal) gpusWanted = atoi(argv[++i]);
        else if ((arg == "-e" || arg == "--engine") && hasVal) {
            string e(argv[++i]);
            if (e == "auto") engine = B200SCAN_ENGINE_AUTO; else if (e == "tensor") engine = B200SCAN_ENGINE_TENSOR;
            else if (e == "gather") engine = B200SCAN_ENGINE_GATHER; else throw runtime_error("Unknown engine: " + e);
        } else if ((arg == "-H" || arg == "--histdir") && hasVal) { histdir = argv[++i]; if (histdir.back() != '/') histdir.push_back('/'); }
        else if ((arg == "-o" || arg == "--output") && hasVal) outputFilename = argv[++i];
        else if (arg == "--stats" && hasVal) { gStats.file = argv[++i]; gTimer.on = true; gTimer.quiet = getenv("BLAMM_B200_TIMING") == nullptr; }
        else { scanUsage(); return EXIT_FAILURE; }
    }
    if (!(absSpec || relSpec || pSpec)) relSpec = true;
    if (absSpec && relSpec) throw runtime_error("Specify either the absolute or relative threshold, not both.");
    if (absSpec && pSpec) throw runtime_error("Specify either the absolute or p-value threshold, not both.");
    if (relSpec && pSpec) throw runtime_error("Specify either the relative or p-value threshold, not both.");

    cout << "Welcome to blamm -- PWM scan module" << endl;
    const double tStart = now();
    Settings settings;
    settings.print();

    SpeciesSet sc;
    sc.loadDict(string(argv[argc - 1]) + ".dict");
    uint64_t chunk = 32ull << 20;
    if (const char* e = getenv("BLAMM_B200_CHUNK")) chunk = max<uint64_t>(strtoull(e, nullptr, 10), 1024);
    // CUDA start-up overlaps the loading of the inputs.  (Showing the driver only the devices a small run can use -- CUDA_VISIBLE_DEVICES
    // set before cuInit -- was tried and removed: on this pool's 8-GPU NVSwitch boxes cuInit took 5.9 s with one device visible
    // against 1.0 s with all eight, profiles/r02_scale8_c3_g1.json.)
    future<int> devCount = async(launch::async, [] { return b200scan_device_count(); });
    MotifSet mc;
    mc.load(argv[argc - 2], true);
    cout << "Loaded " << mc.motifs.size() << " motifs from disk\n";
    cout << "Maximum motif size: " << mc.maxLen() << endl;
    if (mc.motifs.empty()) throw runtime_error("No motifs found in " + string(argv[argc - 2]));
    if (mc.maxLen() > B200SCAN_MAX_MOTIF_LEN)
        throw runtime_error("Motifs longer than " + to_string(B200SCAN_MAX_MOTIF_LEN) + " positions are not supported by this build");
    if (revCompl) { mc.addReverseComplements(); cout << "Scanning both forward and reverse strand of the input sequence(s)" << endl; }
    else cout << "Scanning only the forward strand of the input sequence(s)" << endl;
    cout << "Matrix P has dimensions: " << 4 * mc.maxLen() << " x " << mc.motifs.size() << endl;
    if (absSpec) cout << "Absolute motif score threshold set to: " << absThr << endl;
    else if (relSpec) cout << "Relative motif score threshold set to: " << relThr << endl;
    else cout << "P-value motif score threshold set to: " << pvalue << endl;

    gTimer.add("setup: settings, dict, motifs", now() - tStart);
    const double tDev = now();
    int nDev = devCount.get();
    gTimer.add("setup: wait for the CUDA driver (device count)", now() - tDev);
    if (nDev == 0) throw runtime_error("CUDA error: no sm_100 devices found. Aborting...");
    if (gpusWanted > 0) nDev = min(nDev, gpusWanted);
    cout << "Using " << nDev << " GPU devices" << endl;

    ofstream ofsCutoff("PWMthresholds.txt");
    const int outFd = open(outputFilename.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
    if (outFd < 0) throw runtime_error("Cannot write to file: " + outputFilename);
    struct FdCloser { int fd; ~FdCloser() { if (fd >= 0) close(fd); } } outCloser{outFd};

    const uint64_t halo = mc.maxLen() - 1;
    uint64_t totMatches = 0;
    // one context per device for the whole run, sized for the largest group
    uint64_t maxTot = 1024;
    for (const auto& sp : sc.species) maxTot = max<uint64_t>(maxTot, sp.totSeqLen);
    // size the hit buffers for the expected hit rate (they regrow on demand, at the price of re-scanning a block); where the rate is
    // known to be high (-pt) the chunks shrink so that a chunk's hits stay within ~4e8 records (device memory: ~96 B per hit).
    // Chunks that turn out denser than the device can hold are scored in halves (deviceWorker: scanSplit).
    const double rate = pSpec ? std::min(1.0, 3.0 * pvalue) : 2e-4;
    if (rate * (double)mc.motifs.size() * (double)chunk > 4e8)
        chunk = max<uint64_t>(1 << 20, (uint64_t)(4e8 / (rate * (double)mc.motifs.size())) / 32 * 32);
    const uint64_t maxBlock = min<uint64_t>(chunk, maxTot) + halo + 64;
    const uint64_t maxHits = std::max<uint64_t>(1 << 20, (uint64_t)(rate * (double)maxBlock * (double)mc.motifs.size()));
    struct CtxPool {
        vector<b200scan_ctx*> ctx;
        vector<shared_future<string>> ready;
        ~CtxPool()
        {
            for (auto& r : ready) if (r.valid()) r.wait();
            const double t0 = now();
            for (auto c : ctx) if (c) b200scan_destroy(c);
            gTimer.add("b200scan_destroy (all GPUs)", now() - t0);
        }
    } pool;
    pool.ctx.assign((size_t)nDev, nullptr);
    for (int d = 0; d < nDev; d++)
        pool.ready.push_back(async(launch::async, [&pool, d, maxBlock, maxHits]() -> string {
            const double t0 = now();
            if (b200scan_create(&pool.ctx[(size_t)d], d, maxBlock, maxHits) != B200SCAN_OK) {
                pool.ctx[(size_t)d] = nullptr;
                return string("CUDA error: ") + b200scan_last_error(nullptr);
            }
            gTimer.add("b200scan_create (per GPU)", now() - t0);
            return string();
        }).share());

    // One pipeline for the whole run: this thread plans the groups (P, thresholds) and reads + packs their FASTA files, one worker
    // per GPU scores the chunks, the shared pool formats and writes.  Nothing drains at a group boundary except a GPU's own
    // chunks in flight when it loads the next group's motifs.
    WorkPool workPool(numThreads);
    ScanShared sh;
    for (const auto& m : mc.motifs) sh.mtext.emplace_back(m.name, (uint64_t)m.size(), m.revComp);
    sh.fd = outFd; sh.pool = &workPool;
    sh.chooseWriter();
    sh.maxQueue = (size_t)nDev + 1;
    vector<thread> workers;
    for (int d = 0; d < nDev; d++)
        workers.emplace_back(deviceWorker, ref(sh), engine, foldLower, &pool.ctx[(size_t)d], pool.ready[(size_t)d], (size_t)d, numThreads, halo);
    auto finish = [&](bool failed) {
        { lock_guard<mutex> l(sh.qMutex); sh.done = true; }
        if (failed) sh.fail("input error"); else sh.qCv.notify_all();
        for (auto& w : workers) w.join();
        sh.drainEmitter();                           // the last chunks' text is in the file
    };
    uint64_t nJobs = 0, groupId = 0;
    double tSetup = now();
    try {
        for (const auto& sp : sc.species) {
            if (sh.failed) break;
            cout << "Scanning species: " << sp.name;
            sp.printNuclProb(settings.pseudocount);
            mc.generateMatrix(sp.nuclCounts, settings.pseudocount);
            if (pSpec) {
                // a reverse complement and the permutations of a motif read the histogram of their base name (motif.h:256-267,
                // pwmscan.cpp:607-611): every file is parsed once, on -t threads; the first missing file is reported as before
                vector<string> base; unordered_map<string, size_t> slotOf;
                for (const auto& m : mc.motifs) if (slotOf.emplace(m.baseName(), base.size()).second) base.push_back(m.baseName());
                vector<float> cutoff(base.size()); vector<string> err(base.size());
                workPool.parallel(base.size(), [&](size_t i) {
                    try { ScoreHistogram h; h.load(histdir, "hist_" + sp.name + "_" + base[i]); cutoff[i] = h.scoreCutoff(pvalue); }
                    catch (const exception& e) { err[i] = e.what(); }
                });
                for (const auto& e : err) if (!e.empty()) throw runtime_error(e);
                for (auto& m : mc.motifs) m.threshold = cutoff[slotOf[m.baseName()]];
            } else for (auto& m : mc.motifs) {
                if (absSpec) m.threshold = absThr;
                else { const float mx = m.maxScore(), mn = m.minScore(); m.threshold = relThr * (mx - mn) + mn; }
            }
            for (const auto& m : mc.motifs) {
                if (m.size() > 15) continue;                 // the reference lists only short motifs (pwmscan.cpp:620)
                ofsCutoff << sp.name << "\t" << m.name << "\t" << m.minScore() << "\t" << m.threshold << "\t" << m.maxScore() << "\n";
            }
            auto group = make_shared<GroupParams>();
            group->species = &sp; group->id = groupId++;
            group->P = mc.P(); group->ldp = mc.ldp(); group->len = mc.colLen(); group->thr = mc.colThr();
            size_t sl = 0, ml = 0;
            for (const auto& n : sp.seqNames) sl = max(sl, n.size());
            for (const auto& m : mc.motifs) ml = max(ml, m.name.size());
            group->maxNameLen = sl + ml;
            gTimer.add("setup: matrix P, histograms, thresholds (per group)", now() - tSetup);

            FastaStream fs(sp.files, sp.totSeqLen);
            fs.setParallel(ingestThreads(numThreads));
            const char* asciiEnv = getenv("BLAMM_B200_ASCII");
            const bool sendAscii = asciiEnv && *asciiEnv && *asciiEnv != '0';
            FastaStream::Chunk c;
            for (;;) {
                double tr = now();
                if (sh.failed || !fs.next(maxBlock - halo - 64, halo, c)) break;
                unique_ptr<Job> job(new Job);
                if (sendAscii) {
                    job->chars.reset(new char[c.nTotal]);
                    fs.copyChunk(c, job->chars.get());
                } else {                                  // 0.375 byte per character crosses PCIe instead of 1 (the reference: 18)
                    job->codes.reset(new uint32_t[(c.nTotal + 15) / 16]);
                    job->zmask.reset(new uint32_t[(c.nTotal + 31) / 32]);
                    job->hasZero = fs.packChunk(c, foldLower, job->codes.get(), job->zmask.get());
                }
                job->fragStarts = c.fragStarts; job->frags = c.frags;
                job->nTotal = c.nTotal; job->nPayload = c.nPayload;
                job->seq = nJobs++;
                job->group = group;
                gTimer.add("FASTA read + filter (reader)", now() - tr);
                tr = now();
                unique_lock<mutex> l(sh.qMutex);
                sh.qCv.wait(l, [&] { return sh.queue.size() < sh.maxQueue || sh.failed; });
                sh.queue.push_back(std::move(job));
                sh.qCv.notify_all();
                l.unlock();
                gTimer.add("reader waits for a free queue slot", now() - tr);
                if (sp.totSeqLen) { cout << "Progress... " << (100 * min(fs.filteredLength(), sp.totSeqLen)) / sp.totSeqLen << "%\r"; cout.flush(); }
            }
            cout << "Progress... 100%  " << endl;          // (of the reading: scoring and writing of the group's last chunks go on)
            tSetup = now();
        }
    } catch (...) {
        finish(true);
        throw;
    }
    {
        const double t0 = now();
        finish(false);
        gTimer.add("wait for the last chunks (workers join)", now() - t0);
    }
    if (sh.failed) throw runtime_error(sh.error.empty() ? "scan failed" : sh.error);
    totMatches = sh.totMatches;
    ofsCutoff.close();
    cout << "\nWrote " << totMatches << " matches to " << outputFilename << ".\n";
    gStatsColumns = mc.motifs.size(); gStatsMatches = totMatches;
    gTimer.add("scan module before teardown", now() - tStart);
    return EXIT_SUCCESS;
}

} // namespace blamm

// =========================================================================================================
// main  (reference blstools.cpp:113-166)
// =========================================================================================================
static void usage()
{
    cout << "Usage: blamm command [options]\n\n command\n"
            "  dict\t\t\tmake a dictionary for the input sequences\n"
            "  hist\t\t\tgenerate PWM score histograms\n"
            "  scan\t\t\tscan for pwm occurrences\n\n"
            " [options]\n  -h\t--help\t\tdisplay help page\n  -v\t--version\tdisplay version\n\n";
}

int main(int argc, char** argv)
{
    if (argc < 2) { usage(); return EXIT_FAILURE; }
    const string cmd(argv[1]);
    try {
        if (cmd == "dict") { int rc = blamm::runDict(argc, argv); if (rc == EXIT_SUCCESS) cout << "Exiting... bye!" << endl; return rc; }
        if (cmd == "hist") { int rc = blamm::runHist(argc, argv); if (rc == EXIT_SUCCESS) cout << "Exiting... bye!" << endl; return rc; }
        if (cmd == "selftest-writer") return blamm::runWriterSelfTest(argc, argv);
        if (cmd == "selftest-order") return blamm::runOrderSelfTest(argc, argv);
        if (cmd == "scan") {
            const double t0 = blamm::now();
            int rc = blamm::runScan(argc, argv);
            blamm::gTimer.add("scan module incl. teardown", blamm::now() - t0);
            blamm::gTimer.report();
            blamm::writeStats();
            if (rc == EXIT_SUCCESS) cout << "Exiting... bye!" << endl;
            // every output file is closed and the contexts are destroyed: skip the CUDA runtime's own at-exit teardown (~0.1 s)
            cout.flush(); cerr.flush(); fflush(nullptr);
            _exit(rc);
        }
    } catch (const exception& e) {
        cerr << e.what() << endl;
        return EXIT_FAILURE;
    }
    if (cmd == "-h" || cmd == "--help") { usage(); return EXIT_SUCCESS; }
    if (cmd == "-v" || cmd == "--version") { cout << "blamm-b200 1.0.0 (B200-native scan path; drop-in for blamm 1.0.0)\n"; return EXIT_SUCCESS; }
    usage();
    return EXIT_FAILURE;
}
