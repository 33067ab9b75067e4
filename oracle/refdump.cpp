// TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// refdump -- float-level dump of what the UNMODIFIED reference computes on the `blamm scan` BLAS path.
// It links the reference's own classes (MotifContainer, SpeciesContainer, FastaBatch, SeqMatrix,
// Matrix::sgemm_batch, ScoreHistogram, Settings) from /root/reference/src (see build_ref.sh) and drives
// them the way PWMScan's constructor + scanThreadBLAS do (pwmscan.cpp:528-640, :223-278), because
// PWMScan itself does all its work inside a constructor with private members (pwmscan.h:94-233) and its
// text output keeps only 6 significant digits (pwmscan.cpp:94).  The harness owns no algorithm: every
// number comes out of the reference's code.
//
// usage: refdump <at|rt|pt> <value> <rc 0|1> <histdir|-> <motifs> <manifest> <out.bin>
//   (run from a directory whose settings.cnf, if any, should apply -- same rule as the reference)
//
// out.bin (little endian):
//   "RDMP" u32 nSpecies
//   per species: u32 nameLen, name | u32 nCols | u32 ldp | per col {u32 len, u32 isRC, f32 thr, u32 nameLen, name}
//                | f32 P[ldp*nCols] (column major) | u64 nHits | per hit {u32 seqIdx, u64 seqPos, u32 col, f32 score, f32 naive}
//   score = the BLAS path's R(i,j) (pwmscan.cpp:113); naive = the reference's own Motif::getScore on the same window
//   (motif.cpp:225-239, the `-s` path: plain in-order float adds, case-insensitive).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "motif.h"
#include "species.h"
#include "sequence.h"
#include "settings.h"
#include "matrix.h"

using namespace std;

static void wr(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(1);} }
static void wu32(FILE* f, uint32_t v) { wr(f, &v, 4); }
static void wu64(FILE* f, uint64_t v) { wr(f, &v, 8); }
static void wf32(FILE* f, float v) { wr(f, &v, 4); }
static void wstr(FILE* f, const string& s) { wu32(f, (uint32_t)s.size()); wr(f, s.data(), s.size()); }

struct Hit { uint32_t seqIdx; uint64_t seqPos; uint32_t col; float score; float naive; };

int main(int argc, char** argv)
{
        if (argc != 8) {
                fprintf(stderr, "usage: refdump <at|rt|pt> <value> <rc 0|1> <histdir|-> <motifs> <manifest> <out.bin>\n");
                return 2;
        }
        string mode = argv[1];
        float value = atof(argv[2]);
        bool rc = atoi(argv[3]) != 0;
        string histdir = argv[4];
        if (histdir == "-") histdir = "";
        else if (histdir.back() != '/') histdir.push_back('/');
        string motifFile = argv[5], manifest = argv[6];

        Settings settings;
        SpeciesContainer sc;
        sc.load(manifest + ".dict");
        MotifContainer mc;
        mc.load(motifFile, true);
        if (rc) mc.addReverseComplements();
        mc.generateMatrixTiles(settings.matrix_P_tile_min_zero_area);

        FILE* out = fopen(argv[7], "wb");
        if (!out) { perror("fopen"); return 1; }
        wr(out, "RDMP", 4);
        wu32(out, (uint32_t)sc.size());

        for (auto species : sc) {
                mc.generateMatrix(species.getNuclCounts(), settings.pseudocount);
                for (auto& motif : mc) {
                        if (mode == "at") motif.setThreshold(value);
                        else if (mode == "rt") {
                                float mx = motif.getMaxScore(), mn = motif.getMinScore();
                                float thr = value * (mx - mn) + mn;
                                motif.setThreshold(thr);
                        } else {
                                ScoreHistogram hist;
                                hist.loadHistogram(histdir, "hist_" + species.getName() + "_" + motif.getBaseName());
                                motif.setThreshold(hist.getScoreCutoff(value));
                        }
                }
                const Matrix& P = mc.getMatrix();
                wstr(out, species.getName());
                wu32(out, (uint32_t)P.nCols());
                wu32(out, (uint32_t)P.nRows());
                for (size_t j = 0; j < P.nCols(); j++) {
                        const Motif& m = mc[mc.getMotifIDAtCol(j)];
                        wu32(out, (uint32_t)m.size());
                        wu32(out, m.isRevCompl() ? 1 : 0);
                        wf32(out, m.getThreshold());
                        wstr(out, m.getName());
                }
                wr(out, P.getData(), sizeof(float) * P.nRows() * P.nCols());

                // --- the scan itself, single thread, as scanThreadBLAS drives it ---
                vector<string> filenames = species.getSequenceFilenames();
                FastaBatch seqBatch(filenames, species.getTotalSeqLength());
                size_t overlap = mc.getMaxMotifLen() - 1;
                size_t w = settings.matrix_S_w, h = settings.matrix_S_h;
                SeqMatrix sm(h, w, overlap);
                Matrix R(h, P.nCols());
                const auto tiles = mc.getMatrixTiles();
                SgemmBatchParams p(tiles.size());
                for (size_t i = 0; i < tiles.size(); i++) {
                        p.m[i] = h; p.k[i] = tiles[i].rowEnd; p.n[i] = tiles[i].colEnd - tiles[i].colStart;
                        p.LDA[i] = h; p.LDB[i] = P.nRows(); p.LDC[i] = h;
                        p.alpha[i] = 1.0f; p.beta[i] = 0.0f;
                        p.B_array[i] = P.getData() + tiles[i].colStart * p.LDB[i];
                        p.C_array[i] = R.getData() + tiles[i].colStart * p.LDC[i];
                }
                vector<Hit> hits;
                while (sm.getNextSeqMatrix(seqBatch)) {
                        for (size_t offset = 0; offset < w; offset++) {
                                for (size_t i = 0; i < tiles.size(); i++)
                                        p.A_array[i] = sm.getData() + 4 * offset * p.LDA[i];
                                Matrix::sgemm_batch(p);
                                for (size_t j = 0; j < R.nCols(); j++) {
                                        const Motif& m = mc[mc.getMotifIDAtCol(j)];
                                        const float thr = m.getThreshold();
                                        for (size_t i = 0; i < sm.getNumOccRow(); i++) {
                                                float s = R(i, j);
                                                if (s < thr) continue;
                                                SeqPos sp = sm.getSeqPos(i, offset);
                                                if (m.size() > sm.getRemainingSeqLen(i, offset)) continue;
                                                float naive = m.getScore(sm.block.substr(w * i + offset, m.size()));
                                                hits.push_back(Hit{(uint32_t)sp.getSeqIndex(), (uint64_t)sp.getSeqPos(), (uint32_t)j, s, naive});
                                        }
                                }
                        }
                }
                wu64(out, hits.size());
                for (const Hit& hh : hits) { wu32(out, hh.seqIdx); wu64(out, hh.seqPos); wu32(out, hh.col); wf32(out, hh.score); wf32(out, hh.naive); }
                cerr << "refdump: species " << species.getName() << ": " << P.nRows() << " x " << P.nCols()
                     << ", " << hits.size() << " hits\n";
        }
        fclose(out);
        return 0;
}
