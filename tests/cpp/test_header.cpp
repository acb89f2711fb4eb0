// Exercises include/SequenceAlignment.h the way the reference's own demos use the reference headers
// (reference test/Test.cpp:27-59, include/Test.cpp:95-190): construct an aligner from a ScoringSystem (+ the
// equal<char> functor), call getAlignment, print / compare the three rows.  Known answers are the golden
// vectors captured from the unmodified reference.  Linked against libseqa_cuda.so on the GPU box
// and against the emulator build of the same sources in the CPU-only suite.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "SequenceAlignment.h"

template <typename Ty, Ty Blank> static void rows(AlignedSequence<Ty, Blank> &A, std::string &R1, std::string &R2, std::string &Fl)
{
    R1.clear(); R2.clear(); Fl.clear();
    for (auto &E : A) {
        R1.push_back(E.get(0));
        R2.push_back(E.get(1));
        Fl.push_back(E.match() ? '|' : ' ');
    }
}

static int Failures = 0;
static void expect(const char *What, const std::string &Got, const std::string &Want)
{
    if (Got != Want) {
        std::printf("FAIL %s: got '%s' want '%s'\n", What, Got.c_str(), Want.c_str());
        Failures++;
    }
}
static bool equalChar(char A, char B) { return A == B; }
static bool fuzzyChar(char A, char B) { return A == B || A == 'N' || B == 'N'; }
static char upperOf(char C) { return (C >= 'a' && C <= 'z') ? (char)(C - 32) : C; }
static bool caseless(char A, char B) { return upperOf(A) == upperOf(B); }                         // an equivalence: table-driven
static bool nNever(char A, char B) { return upperOf(A) == upperOf(B) && upperOf(A) != 'N'; }      // 'N' matches nothing, not even 'N'

int main()
{
    std::string S1 = "AAAGAATGCAT", S2 = "AAACTCAT", R1, R2, Fl;
    {
        NeedlemanWunschSA<std::string, char, '-'> SA(ScoringSystem(-1, 2), equalChar); // README.md:31
        AlignedSequence<char, '-'> A = SA.getAlignment(S1, S2);
        rows(A, R1, R2, Fl);
        expect("NW(-1,2) row1", R1, "AAA-GAATGCAT");
        expect("NW(-1,2) row2", R2, "AAAC---T-CAT");
        expect("NW(-1,2) flags", Fl, "|||    | |||");
        std::cout << R1 << "\n" << Fl << "\n" << R2 << "\n";
    }
    {
        HirschbergSA<std::string, char, '-'> SA(ScoringSystem(-1, 2, -1));
        AlignedSequence<char, '-'> A = SA.getAlignment(S1, S2);
        rows(A, R1, R2, Fl);
        expect("Hirschberg(-1,2,-1) row2", R2, "AAA-C-T-CAT"); // differs from NW: reference quirk (SAHirschberg.h:141)
    }
    {
        SmithWatermanSA<std::string, char, '-'> SA(ScoringSystem(-2, 1, -1), equalChar);
        AlignedSequence<char, '-'> A = SA.getAlignment(S1, S2);
        rows(A, R1, R2, Fl);
        expect("SW(-2,1,-1) row1", R1, "AAAGAATG-----CAT");
        expect("SW(-2,1,-1) row2", R2, "--------AAACTCAT");
    }
    {
        GlobalGotohSA<std::string, char, '-'> SA(ScoringSystem(-3, -1, 1, -1, false));
        AlignedSequence<char, '-'> A = SA.getAlignment(S1, S2);
        rows(A, R1, R2, Fl);
        expect("GlobalGotoh row1", R1, "AAA--GAATGCAT");
        expect("GlobalGotoh row2", R2, "AAACT-----CAT");
    }
    {
        LocalGotohSA<std::string, char, '-'> SA(ScoringSystem(-3, -1, 2, -1));
        AlignedSequence<char, '-'> A = SA.getAlignment(S1, S2);
        rows(A, R1, R2, Fl);
        expect("LocalGotoh row1", R1, "AAAG-AATGCAT");
        expect("LocalGotoh row2", R2, "----AAACTCAT");
    }
    {
        MyersMillerSA<std::string, char, '-'> SA(ScoringSystem(-3, -1, 1, -1, false));
        AlignedSequence<char, '-'> A = SA.getAlignment(S1, S2);
        rows(A, R1, R2, Fl);
        expect("MyersMiller row1", R1, "AAAGAA-TGCAT");
        expect("MyersMiller row2", R2, "A---AACT-CAT");
    }
    { // batched entry point + packed accessor
        std::vector<std::pair<std::string, std::string>> Pairs = {{S1, S2}, {"AATCG", "AACG"}, {"", S2}, {S1, ""}};
        SmithWatermanSA<std::string, char, '-'> SW(ScoringSystem(-2, 1, -1, false), equalChar);
        auto All = SW.getAlignments(Pairs);
        rows(All[1], R1, R2, Fl);
        expect("SW batch[1] row1", R1, "AAT--CG"); // include/Test.cpp:36-37 pair
        expect("SW batch[1] row2", R2, "---AACG");
        rows(All[3], R1, R2, Fl);
        expect("SW batch[3] row2 (empty seq2 -> 11 gaps)", R2, "-----------");
        NeedlemanWunschSA<std::string, char, '-'> NW(ScoringSystem(-1, 2));
        seqa::PackedAlignments Pk = NW.getAlignmentsPacked(Pairs);
        if (Pk.size() != 4 || Pk.OpsLen[0] != 12 || Pk.OpsLen[2] != 8 || NW.LastScores.size() != 4) {
            std::printf("FAIL packed accessor\n");
            Failures++;
        }
        // README pair: AAA-GAATGCAT / AAAC---T-CAT  =  3 diag, 1 left, 3 up, 1 diag, 1 up, 3 diag (2-bit wire format)
        const std::vector<std::pair<uint8_t, uint32_t>> Want = {{0, 3}, {2, 1}, {1, 3}, {0, 1}, {1, 1}, {0, 3}};
        if (!Pk.TwoBit || Pk.runs(0) != Want || Pk.op(2, 7) != SEQA_OP_LEFT || Pk.runs(3).size() != 1) {
            std::printf("FAIL packed runs / 2-bit ops\n");
            Failures++;
        }
    }
    { // StaticFuncs::bridgeNW appends NW of a window (reference include/StaticFuncs.h:27-39)
        AlignedSequence<char, '-'> Res;
        StaticFuncs<std::string, char, '-'>::bridgeNW(S1, S2, Res, ScoringSystem(-1, 2), 3, 3, 11, 8, nullptr);
        rows(Res, R1, R2, Fl);
        NeedlemanWunschSA<std::string, char, '-'> NW(ScoringSystem(-1, 2));
        std::string W1 = S1.substr(3), W2 = S2.substr(3), Q1, Q2, QF;
        AlignedSequence<char, '-'> Direct = NW.getAlignment(W1, W2);
        rows(Direct, Q1, Q2, QF);
        expect("bridgeNW row1", R1, Q1);
        expect("bridgeNW row2", R2, Q2);
    }
    { // StaticFuncs::bridgeNWBatch: the windows between anchors of a chaining aligner as ONE GPU batch
        typedef StaticFuncs<std::string, char, '-'> SF;
        std::string G1 = "ACGTACGTTTGACCAGTACGATCGATCGGCTA", G2 = "ACGTACGTTGACCAGTCGATCGGATCGGTTA";
        std::vector<SF::Window> Ws = {{0, 0, 8, 8}, {8, 8, 16, 15}, {16, 15, 32, 31}, {3, 3, 3, 9}};
        std::vector<AlignedSequence<char, '-'>> Batch;
        SF::bridgeNWBatch(G1, G2, Ws, Batch, ScoringSystem(-1, 2, -1), nullptr);
        for (size_t K = 0; K < Ws.size(); K++) {
            AlignedSequence<char, '-'> One;
            SF::bridgeNW(G1, G2, One, ScoringSystem(-1, 2, -1), Ws[K].Idx1, Ws[K].Idx2, Ws[K].EndIdx1, Ws[K].EndIdx2, nullptr);
            std::string B1, B2, BF;
            rows(Batch[K], B1, B2, BF);
            rows(One, R1, R2, Fl);
            expect("bridgeNWBatch row1", B1, R1);
            expect("bridgeNWBatch row2", B2, R2);
            expect("bridgeNWBatch flags", BF, Fl);
        }
    }
    { // default-constructed affine aligners are defined here: GapOpen = -1, GapExtend = -1, Match = 2, Mismatch = -1
        GlobalGotohSA<std::string, char, '-'> D;
        GlobalGotohSA<std::string, char, '-'> E(ScoringSystem(-1, -1, 2, -1));
        AlignedSequence<char, '-'> A = D.getAlignment(S1, S2), B = E.getAlignment(S1, S2);
        std::string Q1, Q2, QF;
        rows(A, R1, R2, Fl);
        rows(B, Q1, Q2, QF);
        expect("default GlobalGotoh row1", R1, Q1);
        expect("default GlobalGotoh row2", R2, Q2);
        LocalGotohSA<std::string, char, '-'> DL;
        MyersMillerSA<std::string, char, '-'> DM;
        if (DL.getAlignment(S1, S2).size() < S1.size() || DM.getAlignment(S1, S2).size() < S1.size()) {
            std::printf("FAIL default-constructed LocalGotoh / MyersMiller\n");
            Failures++;
        }
    }
    { // LocalGotoh shapes that are undefined behaviour in the reference (SALocalGotoh.h:484-488): rejected per pair
        std::string U1(60, 'A'), U2(57, 'C');
        LocalGotohSA<std::string, char, '-'> LG(ScoringSystem(-3, -1, 1, -1));
        bool Threw = false;
        try {
            LG.getAlignment(U1, U2);
        } catch (const std::invalid_argument &) {
            Threw = true;
        }
        std::vector<std::pair<std::string, std::string>> Pairs = {{S1, S2}, {U1, U2}, {S1, S2}};
        auto All = LG.getAlignments(Pairs);
        rows(All[2], R1, R2, Fl);
        if (!Threw || All.size() != 3 || All[1].size() != 0 || LG.LastUnsupported != std::vector<size_t>{1} || All[0].size() != All[2].size()) {
            std::printf("FAIL per-pair rejection of LocalGotoh UB shapes\n");
            Failures++;
        }
        expect("LocalGotoh batch beside a rejected pair", R2, "--------AAACTCAT");
    }
    { // input wire format: 2-bit symbols when everything is ACGT, 8-bit otherwise -- same alignments either way
        std::vector<std::pair<std::string, std::string>> Clean, Dirty;
        for (int K = 0; K < 40; K++) {
            std::string A, B;
            for (int I = 0; I < 5 + 7 * K % 90; I++) A.push_back("ACGT"[(I * 7 + K * 3 + I / 5) & 3]);
            for (int I = 0; I < 3 + 11 * K % 70; I++) B.push_back("ACGT"[(I * 5 + K + I / 3) & 3]);
            Clean.emplace_back(A, B);
        }
        Dirty = Clean;
        Dirty[7].first[2] = 'N';
        NeedlemanWunschSA<std::string, char, '-'> NW(ScoringSystem(-1, 2, -1));
        auto Two = NW.getAlignments(Clean);
        const bool WasTwo = NW.LastInputsTwoBit;
        NW.ForceByteInputs = true;
        auto Byte = NW.getAlignments(Clean);
        const bool WasByte = !NW.LastInputsTwoBit;
        NW.ForceByteInputs = false;
        NW.getAlignments(Dirty);
        if (!WasTwo || !WasByte || NW.LastInputsTwoBit) {
            std::printf("FAIL input wire format selection\n");
            Failures++;
        }
        for (size_t K = 0; K < Clean.size(); K++) {
            std::string Q1, Q2, QF;
            rows(Two[K], R1, R2, Fl);
            rows(Byte[K], Q1, Q2, QF);
            expect("2-bit vs 8-bit inputs row1", R1, Q1);
            expect("2-bit vs 8-bit inputs row2", R2, Q2);
            expect("2-bit vs 8-bit inputs flags", Fl, QF);
        }

    }
    { // table-driven equality: functors that are an equivalence on bytes run on the GPU path
        std::string L1 = "aaagaATGCat", L2 = "AAACtcAT", U1 = "AAAGAATGCAT", U2 = "AAACTCAT";
        NeedlemanWunschSA<std::string, char, '-'> CL(ScoringSystem(-1, 2), caseless), EQ(ScoringSystem(-1, 2));
        AlignedSequence<char, '-'> A = CL.getAlignment(L1, L2), B = EQ.getAlignment(U1, U2);
        std::string Q1, Q2, QF;
        rows(A, R1, R2, Fl);
        rows(B, Q1, Q2, QF);
        expect("caseless flags", Fl, QF);
        expect("caseless row1 keeps the caller's symbols", R1, "aaa-gaATGCat");
        expect("caseless row2", R2, "AAAC---t-cAT");
        // 'N' matches nothing: an N column scores as a mismatch even against another N
        std::string N1 = "ACGTNNACGT", N2 = "ACGTNNACGT";
        SmithWatermanSA<std::string, char, '-'> NN(ScoringSystem(-2, 1, -1), nNever);
        AlignedSequence<char, '-'> C = NN.getAlignment(N1, N2);
        rows(C, R1, R2, Fl);
        expect("N-never flags", Fl, "||||  ||||");
        if (NN.LastScores.empty() || NN.LastScores[0] != 6) { // 4 + 4 - 1 - 1
            std::printf("FAIL N-never score %d\n", NN.LastScores.empty() ? -999 : NN.LastScores[0]);
            Failures++;
        }
    }
    { // a functor that is not equality is outside the GPU path
        bool Threw = false;
        try {
            NeedlemanWunschSA<std::string, char, '-'> SA(ScoringSystem(-1, 2), fuzzyChar);
            SA.getAlignment(S1, S2);
        } catch (const std::invalid_argument &) {
            Threw = true;
        }
        if (!Threw) {
            std::printf("FAIL custom functor accepted\n");
            Failures++;
        }
    }
    if (Failures) return 1;
    std::printf("OK\n");
    return 0;
}
