// StaticFuncs::bridgeNWBatch against the ORACLE: the windows an anchor-chaining aligner
// (reference include/SAMummer.h:49,75,101, include/SABLAT.h:399) hands to bridgeNW -- the stretches between
// consecutive exact-match anchors of two related sequences -- submitted as one batch through the header, every window
// compared op by op with oracle_align("nw") (oracle/seqa_oracle.c, test infrastructure) on the two substrings.
//   test_bridge_oracle  -> prints "OK <windows> windows" or FAIL lines
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "SequenceAlignment.h"

extern "C" int oracle_align(int algo, int gap, int gap_open, int gap_extend, int match, int mismatch, int allow, const char *s1, int l1,
                            const char *s2, int l2, unsigned char *ops, int cap, int *meta);

static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main()
{
    typedef StaticFuncs<std::string, char, '-'> SF;
    int Failures = 0, Total = 0;
    for (int Trial = 0; Trial < 6; Trial++) {
        // sequence 2 = sequence 1 with substitutions / insertions / deletions; anchors = maximal runs of >= K copied symbols
        const int L = Trial < 3 ? 3000 : 12000, K = 12;
        std::string A, B;
        std::vector<int> Src; // for every symbol of B that is an unmodified copy: its position in A, else -1
        uint64_t Z = 1234 + Trial;
        for (int I = 0; I < L; I++) {
            Z = splitmix64(Z);
            A.push_back("ACGT"[Z & 3]);
            const unsigned U = (Z >> 8) % 100;
            if (U < 3) continue;                                                       // deletion
            if (U < 6) { B.push_back("ACGT"[(Z >> 20) & 3]); Src.push_back(-1); }      // insertion before the copy
            if (U < 14 && U >= 6) { B.push_back("ACGT"[((Z & 3) + 1 + (Z >> 24) % 3) & 3]); Src.push_back(-1); continue; } // substitution
            B.push_back(A.back());
            Src.push_back(I);
        }
        struct Anchor { int I, J, Len; };
        std::vector<Anchor> Anchors;
        for (size_t J = 0; J < B.size();) {
            if (Src[J] < 0) { J++; continue; }
            size_t E = J + 1;
            while (E < B.size() && Src[E] >= 0 && Src[E] == Src[E - 1] + 1) E++;
            if ((int)(E - J) >= K) Anchors.push_back({Src[J], (int)J, (int)(E - J)});
            J = E;
        }
        // the windows between anchors (and before the first / after the last), as SAMummer.h:49,75,101 cuts them
        std::vector<SF::Window> Ws;
        int PI = 0, PJ = 0;
        for (const Anchor &An : Anchors) {
            Ws.push_back({PI, PJ, An.I, An.J});
            PI = An.I + An.Len;
            PJ = An.J + An.Len;
        }
        Ws.push_back({PI, PJ, (int)A.size(), (int)B.size()});
        const ScoringSystem Sc = (Trial & 1) ? ScoringSystem(-1, 2) : ScoringSystem(-1, 2, -1);
        std::vector<AlignedSequence<char, '-'>> Parts;
        SF::bridgeNWBatch(A, B, Ws, Parts, Sc, nullptr);
        for (size_t W = 0; W < Ws.size(); W++) {
            const int L1 = Ws[W].EndIdx1 - Ws[W].Idx1, L2 = Ws[W].EndIdx2 - Ws[W].Idx2;
            std::vector<unsigned char> Want((size_t)(L1 + L2 + 8));
            int Meta[5];
            const int NOps = oracle_align(0, Sc.getGapPenalty(), 0, 0, Sc.getMatchProfit(), Sc.getMismatchPenalty(), Sc.getAllowMismatch() ? 1 : 0,
                                          A.data() + Ws[W].Idx1, L1, B.data() + Ws[W].Idx2, L2, Want.data(), (int)Want.size(), Meta);
            std::vector<unsigned char> Got;
            int I = Ws[W].Idx1, J = Ws[W].Idx2;
            bool SymOk = true;
            for (auto &E : Parts[W]) {
                if (E.get(0) != '-' && E.get(1) != '-') { SymOk &= E.get(0) == A[I] && E.get(1) == B[J] && E.match() == (A[I] == B[J]); I++; J++; Got.push_back(0); }
                else if (E.get(1) == '-') { SymOk &= E.get(0) == A[I]; I++; Got.push_back(1); }
                else { SymOk &= E.get(1) == B[J]; J++; Got.push_back(2); }
            }
            Want.resize((size_t)(NOps < 0 ? 0 : NOps));
            if (NOps < 0 || Got != Want || !SymOk || I != Ws[W].EndIdx1 || J != Ws[W].EndIdx2) {
                std::printf("FAIL trial %d window %zu (%d x %d)\n", Trial, W, L1, L2);
                Failures++;
            }
            Total++;
        }
    }
    if (Failures) return 1;
    std::printf("OK %d windows\n", Total);
    return 0;
}
