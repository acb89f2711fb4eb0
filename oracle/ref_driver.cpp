// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin C-ABI driver around the UNMODIFIED reference headers, which are included
// BY PATH from /root/reference/include at build time (see oracle/Makefile); no
// reference source is copied into this repository.  The resulting shared object
// lives in oracle/_ref/ (git-ignored, travels to the GPU box with gpurun) and is
// used (a) to pin oracle/seqa_oracle.c, (b) to generate tests/golden/*.json and
// (c) as bench.py's `--impl reference` / cpu_baseline "reference" arm.
//
// Calling rules follow SURVEY.md section 8c: SmithWatermanSA is always given an
// explicit equality functor (its nullptr path reads an uninitialised member,
// reference include/SASmithWaterman.h:53-54), affine aligners always get the
// 5-argument ScoringSystem, one fresh aligner object per call / per thread.
#include <limits>
#include <cmath>
#include <vector>
#include <map>
#include <unordered_map>
#include <string>
#include <chrono>
#include <queue>
#include <thread>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include <functional>
#include <list>
#include <iostream>
#include <fstream>
#include <cassert>
#include <limits.h>

// Score access (SURVEY.md 8c "Getting scores"): expose the private matrices of the
// reference classes to this translation unit only.
#define private public
#include "SequenceAlignment.h"
#undef private

static bool eq_char(char a, char b) { return a == b; }

typedef AlignedSequence<char, '-'> Aln;

static ScoringSystem make_scoring(int algo, int ctor, int gap, int go, int ge, int match, int mismatch, int allow)
{
    // ctor: 2 -> ScoringSystem(Gap,Match); 4 -> (Gap,Match,Mismatch,Allow); 5 -> (GapOpen,GapExtend,Match,Mismatch,Allow)
    (void)algo;
    if (ctor == 2) return ScoringSystem(gap, match);
    if (ctor == 4) return ScoringSystem(gap, match, mismatch, allow != 0);
    return ScoringSystem(go, ge, match, mismatch, allow != 0);
}

static Aln run_one(int algo, ScoringSystem sc, int use_functor, std::string &s1, std::string &s2,
                   int *score, int *max_row, int *max_col)
{
    std::function<bool(char, char)> fn = nullptr;
    if (use_functor) fn = eq_char;
    if (score) *score = INT_MIN;
    if (max_row) *max_row = -1;
    if (max_col) *max_col = -1;
    switch (algo)
    {
    case 0:
    {
        NeedlemanWunschSA<std::string, char, '-'> SA(sc, fn);
        if (score)
        {
            SA.cacheAllMatches(s1, s2);
            SA.computeScoreMatrix(s1, s2);
            *score = SA.Matrix[(s1.size() + 1) * (s2.size() + 1) - 1];
            SA.clearAll();
        }
        return SA.getAlignment(s1, s2);
    }
    case 1:
    {
        SmithWatermanSA<std::string, char, '-'> SA(sc, eq_char); // functor is mandatory (see header comment)
        Aln r = SA.getAlignment(s1, s2);
        if (score && s1.size() > 0 && s2.size() > 0)
        {
            *score = SA.MaxScore;
            *max_row = (int)SA.MaxRow;
            *max_col = (int)SA.MaxCol;
        }
        return r;
    }
    case 2:
    {
        GlobalGotohSA<std::string, char, '-'> SA(sc, fn);
        if (score)
        {
            SA.cacheAllMatches(s1, s2);
            SA.computeScoreMatrix(s1, s2);
            *score = SA.Matrix[(s1.size() + 1) * (s2.size() + 1) - 1];
            SA.clearAll();
        }
        return SA.getAlignment(s1, s2);
    }
    case 3:
    {
        LocalGotohSA<std::string, char, '-'> SA(sc, fn);
        if (score && s1.size() > 0 && s2.size() > 0)
        {
            SA.cacheAllMatches(s1, s2);
            SA.computeScoreMatrix(s1, s2);
            *score = SA.Matrix[SA.MaxRow * (s2.size() + 1) + SA.MaxCol];
            *max_row = (int)SA.MaxRow;
            *max_col = (int)SA.MaxCol;
            SA.clearAll();
        }
        return SA.getAlignment(s1, s2);
    }
    case 4:
    {
        HirschbergSA<std::string, char, '-'> SA(sc, fn);
        return SA.getAlignment(s1, s2);
    }
    default:
    {
        MyersMillerSA<std::string, char, '-'> SA(sc, fn);
        return SA.getAlignment(s1, s2);
    }
    }
}

extern "C" {

// Returns the number of alignment columns (entries); row1/row2/flags receive up to `cap` of them.
// score/max_row/max_col may be NULL.  score == INT_MIN when the reference exposes none (algo 4, 5).
int ref_align(int algo, int ctor, int gap, int gap_open, int gap_extend, int match, int mismatch, int allow_mismatch,
              int use_functor, const char *seq1, int n1, const char *seq2, int n2,
              char *row1, char *row2, unsigned char *flags, int cap,
              int *score, int *max_row, int *max_col)
{
    std::string s1(seq1, seq1 + n1), s2(seq2, seq2 + n2);
    ScoringSystem sc = make_scoring(algo, ctor, gap, gap_open, gap_extend, match, mismatch, allow_mismatch);
    Aln r = run_one(algo, sc, use_functor, s1, s2, score, max_row, max_col);
    int k = 0;
    for (auto &e : r)
    {
        if (k < cap)
        {
            row1[k] = e.get(0);
            row2[k] = e.get(1);
            flags[k] = e.match() ? 1 : 0;
        }
        k++;
    }
    return k;
}

// Threaded throughput run of the reference's own getAlignment over a batch:
// one aligner object per std::thread, contiguous static split.  Returns wall seconds; *entries_out
// receives the total number of alignment columns produced (keeps the work observable).
double ref_bench(int algo, int ctor, int gap, int gap_open, int gap_extend, int match, int mismatch, int allow_mismatch,
                 const char *bases, const uint64_t *off1, const uint64_t *off2,
                 const uint32_t *len1, const uint32_t *len2, uint64_t n_pairs, int threads,
                 uint64_t *entries_out)
{
    if (threads < 1) threads = 1;
    std::vector<uint64_t> cnt(threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
    {
        th.emplace_back([&, t]() {
            uint64_t lo = n_pairs * t / threads, hi = n_pairs * (t + 1) / threads;
            ScoringSystem sc = make_scoring(algo, ctor, gap, gap_open, gap_extend, match, mismatch, allow_mismatch);
            std::function<bool(char, char)> nofn = nullptr;
            NeedlemanWunschSA<std::string, char, '-'> nw(sc, nofn);
            SmithWatermanSA<std::string, char, '-'> sw(sc, eq_char);
            GlobalGotohSA<std::string, char, '-'> gg(sc, nofn);
            LocalGotohSA<std::string, char, '-'> lg(sc, nofn);
            HirschbergSA<std::string, char, '-'> hb(sc, nofn);
            MyersMillerSA<std::string, char, '-'> mm(sc, nofn);
            uint64_t c = 0;
            for (uint64_t p = lo; p < hi; p++)
            {
                std::string s1(bases + off1[p], bases + off1[p] + len1[p]);
                std::string s2(bases + off2[p], bases + off2[p] + len2[p]);
                // note: the reference's AlignedSequence::operator= does not compile for Blank != 0
                // (include/SequenceAlignment.h:62-66), so results are only ever move-constructed here.
                switch (algo)
                {
                case 0: { Aln r(nw.getAlignment(s1, s2)); c += r.Data.size(); break; }
                case 1: { Aln r(sw.getAlignment(s1, s2)); c += r.Data.size(); break; }
                case 2: { Aln r(gg.getAlignment(s1, s2)); c += r.Data.size(); break; }
                case 3: { Aln r(lg.getAlignment(s1, s2)); c += r.Data.size(); break; }
                case 4: { Aln r(hb.getAlignment(s1, s2)); c += r.Data.size(); break; }
                default: { Aln r(mm.getAlignment(s1, s2)); c += r.Data.size(); break; }
                }
            }
            cnt[t] = c;
        });
    }
    for (auto &x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    uint64_t tot = 0;
    for (auto c : cnt) tot += c;
    if (entries_out) *entries_out = tot;
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
}
