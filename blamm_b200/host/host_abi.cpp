// C ABI over the host model (include/blamm_host.h).
#include "host.h"
#include "../../include/blamm_host.h"

#include <cstring>

using namespace blamm;

struct blamm_motifs { MotifSet set; };
struct blamm_fasta {
    std::unique_ptr<FastaStream> fs;
    FastaStream::Chunk chunk;
    std::vector<uint64_t> fstart, fseq, fpos;
};

static thread_local std::string g_err;
template <class F> static int guarded(F&& f)
{
    try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

extern "C" {

const char* blamm_host_last_error(void) { return g_err.c_str(); }

int blamm_motifs_load(const char* path, int loadPerm, int addRc, blamm_motifs** out)
{
    return guarded([&] {
        std::unique_ptr<blamm_motifs> m(new blamm_motifs);
        m->set.load(path, loadPerm != 0);
        if (addRc) m->set.addReverseComplements();
        *out = m.release();
    });
}
void blamm_motifs_free(blamm_motifs* m) { delete m; }
int blamm_motifs_count(const blamm_motifs* m) { return (int)m->set.motifs.size(); }
int blamm_motifs_max_len(const blamm_motifs* m) { return (int)m->set.maxLen(); }

int blamm_motifs_generate_matrix(blamm_motifs* m, const uint64_t bg[4], float pseudo)
{
    return guarded([&] { m->set.generateMatrix({bg[0], bg[1], bg[2], bg[3]}, pseudo); });
}

int blamm_motifs_get_matrix(const blamm_motifs* m, float* P, uint64_t n)
{
    return guarded([&] {
        if (n != m->set.P().size()) throw std::runtime_error("blamm_motifs_get_matrix: size mismatch");
        std::memcpy(P, m->set.P().data(), n * sizeof(float));
    });
}

int blamm_motifs_get_columns(const blamm_motifs* m, int32_t* len, uint8_t* rc, float* mn, float* mx)
{
    return guarded([&] {
        for (size_t c = 0; c < m->set.motifs.size(); c++) {
            const Motif& mo = m->set.motifs[c];
            if (len) len[c] = (int32_t)mo.size();
            if (rc) rc[c] = mo.revComp;
            if (mn) mn[c] = mo.minScore();
            if (mx) mx[c] = mo.maxScore();
        }
    });
}

const char* blamm_motifs_name(const blamm_motifs* m, int col) { return m->set.motifs.at(col).name.c_str(); }

int blamm_motifs_set_thresholds(blamm_motifs* m, int mode, float value, const char* species, const char* histdir, float* out)
{
    return guarded([&] {
        std::string dir = histdir ? histdir : "";
        if (!dir.empty() && dir.back() != '/') dir.push_back('/');
        size_t c = 0;
        for (auto& mo : m->set.motifs) {
            if (mode == 0) mo.threshold = value;
            else if (mode == 1) { const float mx = mo.maxScore(), mn = mo.minScore(); mo.threshold = value * (mx - mn) + mn; }
            else if (mode == 2) { ScoreHistogram h; h.load(dir, std::string("hist_") + species + "_" + mo.baseName()); mo.threshold = h.scoreCutoff(value); }
            else throw std::runtime_error("unknown threshold mode");
            if (out) out[c] = mo.threshold;
            c++;
        }
    });
}

int blamm_motifs_write_histograms(blamm_motifs* m, const uint64_t bg[4], float pseudo, uint64_t bins, uint64_t maxLength,
                                  const char* species, const char* histdir)
{
    return guarded([&] {
        std::string dir = histdir ? histdir : "";
        if (!dir.empty() && dir.back() != '/') dir.push_back('/');
        Species sp; sp.name = species; sp.nuclCounts = {bg[0], bg[1], bg[2], bg[3]};
        m->set.generateMatrix(sp.nuclCounts, pseudo);
        const auto prob = sp.nuclProb(pseudo);
        for (const auto& mo : m->set.motifs) {
            if (mo.revComp) continue;
            ScoreHistogram h(mo.minScore(), mo.maxScore(), bins);
            MotifSet::theoreticalHistogram(mo, prob, bins, maxLength, h);
            h.writeGNUPlot(dir, "hist_" + sp.name + "_" + mo.name, mo.name + " (" + sp.name + ")");
        }
    });
}

int blamm_fasta_open(const char* const* files, int n, uint64_t maxFiltered, blamm_fasta** out)
{
    return guarded([&] {
        std::vector<std::string> v(files, files + n);
        std::unique_ptr<blamm_fasta> f(new blamm_fasta);
        f->fs.reset(new FastaStream(v, maxFiltered));
        *out = f.release();
    });
}
void blamm_fasta_close(blamm_fasta* f) { delete f; }
int blamm_fasta_set_parallel(blamm_fasta* f, unsigned threads, uint64_t segmentBytes)
{
    return guarded([&] { f->fs->setParallel(threads, (size_t)segmentBytes); });
}

int blamm_fasta_next(blamm_fasta* f, uint64_t payload, uint64_t halo, const char** chars, uint64_t* nTotal, uint64_t* nPayload,
                     uint64_t* streamStart, const uint64_t** fs, const uint64_t** fq, const uint64_t** fp, uint64_t* nFrag)
{
    int more = 0;
    int rc = guarded([&] {
        more = f->fs->next(payload, halo, f->chunk) ? 1 : 0;
        f->fstart.clear(); f->fseq.clear(); f->fpos.clear();
        for (const auto& fr : f->chunk.frags) { f->fstart.push_back(fr.streamPos); f->fseq.push_back(fr.seqIdx); f->fpos.push_back(fr.seqPos); }
        *chars = f->chunk.chars; *nTotal = f->chunk.nTotal; *nPayload = f->chunk.nPayload; *streamStart = f->chunk.streamStart;
        *fs = f->fstart.data(); *fq = f->fseq.data(); *fp = f->fpos.data(); *nFrag = f->fstart.size();
    });
    return rc ? rc : more;
}

int blamm_fasta_num_sequences(const blamm_fasta* f) { return (int)f->fs->seqNames().size(); }
const char* blamm_fasta_sequence_name(const blamm_fasta* f, int idx) { return f->fs->seqNames().at(idx).c_str(); }
int blamm_fasta_counts(const blamm_fasta* f, uint64_t c[4]) { for (int i = 0; i < 4; i++) c[i] = f->fs->counts()[i]; return 0; }

int blamm_pack_ascii(const char* chars, uint64_t n, int foldLower, uint32_t* codes2, uint32_t* zeroMask)
{
    int any = 0;
    int rc = guarded([&] {
        if ((!chars && n) || !codes2 || !zeroMask) throw std::runtime_error("blamm_pack_ascii: NULL argument");
        any = packAscii(chars, n, foldLower != 0, codes2, zeroMask) ? 1 : 0;
    });
    return rc ? -1 : any;
}
int blamm_fasta_pack(blamm_fasta* f, int foldLower, uint32_t* codes2, uint32_t* zeroMask)
{
    int any = 0;
    int rc = guarded([&] {
        if (!f || !codes2 || !zeroMask) throw std::runtime_error("blamm_fasta_pack: NULL argument");
        if (!f->chunk.chars) throw std::runtime_error("blamm_fasta_pack: no chunk (call blamm_fasta_next first)");
        any = f->fs->packChunk(f->chunk, foldLower != 0, codes2, zeroMask) ? 1 : 0;
    });
    return rc ? -1 : any;
}

int blamm_format_score(float score, char* dst) { return dst ? formatScore(dst, score) : -1; }

} // extern "C"
