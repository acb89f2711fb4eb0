// SequenceAlignment.h -- host-side mirror of SeqALib's header-only C++ API for its DP hot path, backed by the
// B200 library libseqa_cuda.so through the C ABI in seqa_cuda.h.
//
// Same class templates, constructors and call shapes as the reference (przemektmalon/SeqALib):
//   AlignedSequence<Ty,Blank>, ScoringSystem          reference include/SequenceAlignment.h:13-131
//   SequenceAligner<ContainerType,Ty,Blank,MatchFnTy>  reference include/SequenceAlignment.h:133-189
//   NeedlemanWunschSA / SmithWatermanSA / GlobalGotohSA / LocalGotohSA / HirschbergSA / MyersMillerSA
//                                                      reference include/SA*.h (one class per file there)
//   ArrayView<ContainerType>, StaticFuncs::useNW / bridgeNW
//                                                      reference include/ArrayView.h:4-54, include/StaticFuncs.h:6-40
// plus the batched entry point the reference lacks:
//   std::vector<AlignedSequence<Ty,Blank>> getAlignments(std::vector<std::pair<ContainerType,ContainerType>>&)
//   seqa::PackedAlignments getAlignmentsPacked(...)   (scores + op strings, no std::list materialisation)
//
// What runs where: the DP fill, the traceback and the Hirschberg / Myers-Miller recursion run on the GPU; this
// header only packs the inputs, calls seqa_cuda_align_batch and expands the returned op strings (0 diagonal,
// 1 up, 2 left) into AlignedSequence entries, adding the reference's forceGlobal framing for the two local
// algorithms.  There is no CPU implementation of the algorithms here.  The GPU path takes 8-bit symbols compared
// with ==; a user matching functor is accepted when it IS equality on the whole 8-bit domain (checked
// exhaustively once per aligner: 65,536 calls) -- that is what the reference's own demos pass
// (test/Test.cpp:44) and what SmithWatermanSA needs in the reference to be defined behaviour.  Any other functor,
// or a wider symbol type, is outside this path (std::invalid_argument): use the reference implementation there.
// Errors of the device path surface as std::runtime_error carrying seqa_cuda_last_error().
#ifndef SEQA_SEQUENCE_ALIGNMENT_H
#define SEQA_SEQUENCE_ALIGNMENT_H

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <list>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "seqa_cuda.h"

#if defined(__GNUC__) && defined(__x86_64__)
#include <immintrin.h>
#define SEQA_HAVE_AVX2_DISPATCH 1
#endif

#define ScoreSystemType int

template <typename Ty, Ty Blank = Ty(0)> class AlignedSequence {
  public:
    class Entry {
        Ty First, Second;
        bool Matching;

      public:
        Entry() : First(Blank), Second(Blank), Matching(false) {}
        Entry(Ty V1, Ty V2) : First(V1), Second(V2), Matching(V1 != Blank && V2 != Blank) {}
        Entry(Ty V1, Ty V2, bool IsMatch) : First(V1), Second(V2), Matching(IsMatch) {}
        Ty get(size_t Index) const { return Index == 0 ? First : Second; }
        bool empty() const { return First == Blank && Second == Blank; }
        bool hasBlank() const { return First == Blank || Second == Blank; }
        bool match() const { return Matching; }
        bool mismatch() const { return !Matching; }
        Ty getNonBlank() const { return First != Blank ? First : Second; }
    };

    std::list<Entry> Data;

    AlignedSequence() = default;
    AlignedSequence(const AlignedSequence &) = default;
    AlignedSequence(AlignedSequence &&) = default;
    AlignedSequence &operator=(const AlignedSequence &) = default;
    AlignedSequence &operator=(AlignedSequence &&) = default;

    void append(const AlignedSequence &Other) { Data.insert(Data.end(), Other.Data.begin(), Other.Data.end()); }
    void splice(AlignedSequence &Other) { Data.splice(Data.end(), Other.Data); }
    typename std::list<Entry>::iterator begin() { return Data.begin(); }
    typename std::list<Entry>::iterator end() { return Data.end(); }
    size_t size() const { return Data.size(); }
};

// Same three constructors as the reference.  Unlike the reference, members the chosen constructor does not
// set are zero instead of indeterminate.
class ScoringSystem {
    ScoreSystemType Gap = 0, Match = 0, Mismatch = 0, GapOpen = 0, GapExtend = 0;
    bool AllowMismatch = false;

  public:
    ScoringSystem(ScoreSystemType Gap, ScoreSystemType Match)
        : Gap(Gap), Match(Match), Mismatch(std::numeric_limits<ScoreSystemType>::min()), AllowMismatch(false) {}
    ScoringSystem(ScoreSystemType Gap, ScoreSystemType Match, ScoreSystemType Mismatch, bool AllowMismatch = true)
        : Gap(Gap), Match(Match), Mismatch(Mismatch), AllowMismatch(AllowMismatch) {}
    ScoringSystem(ScoreSystemType GapOpen, ScoreSystemType GapExtend, ScoreSystemType Match, ScoreSystemType Mismatch,
                  bool AllowMismatch = true)
        : Match(Match), Mismatch(Mismatch), GapOpen(GapOpen), GapExtend(GapExtend), AllowMismatch(AllowMismatch) {}

    bool getAllowMismatch() const { return AllowMismatch; }
    ScoreSystemType getMismatchPenalty() const { return Mismatch; }
    ScoreSystemType getGapPenalty() const { return Gap; }
    ScoreSystemType getMatchProfit() const { return Match; }
    ScoreSystemType getGapOpenPenalty() const { return GapOpen; }
    ScoreSystemType getGapExtendPenalty() const { return GapExtend; }
};

// Non-owning window over a contiguous container (reference include/ArrayView.h:4-54).
template <typename ContainerType> class ArrayView {
  public:
    using value_type = typename ContainerType::value_type;
    using iterator = const value_type *;

  private:
    const value_type *Base = nullptr;
    size_t Lo = 0, Hi = 0;

  public:
    ArrayView() = default;
    ArrayView(const ContainerType &C) : Base(C.size() ? &*C.begin() : nullptr), Lo(0), Hi(C.size()) {}
    // window [Start, End) RELATIVE to the current window
    void sliceWindow(size_t Start, size_t End)
    {
        Hi = Lo + End;
        Lo = Lo + Start;
    }
    size_t size() const { return Hi - Lo; }
    const value_type &operator[](size_t I) const { return Base[Lo + I]; }
    iterator begin() const { return Base + Lo; }
    iterator end() const { return Base + Hi; }
    const value_type *data() const { return Base + Lo; }
};

namespace seqa {

// Raw results of a batch: what the C ABI returns, owned by vectors.
// Grow-only block of page-locked host memory from the library (seqa_cuda_host_alloc): batch buffers in such memory
// cross PCIe at full speed and are not page-faulted in on every call.
struct PinnedBlock {
    char *P = nullptr;
    size_t Cap = 0;
    PinnedBlock() = default;
    PinnedBlock(const PinnedBlock &) = delete;
    PinnedBlock &operator=(const PinnedBlock &) = delete;
    ~PinnedBlock()
    {
        if (P) seqa_cuda_host_free(P);
    }
    void reserve(size_t Bytes)
    {
        if (Bytes <= Cap) return;
        if (P) seqa_cuda_host_free(P);
        P = nullptr;
        Cap = 0;
        const size_t Want = Bytes + Bytes / 8 + 4096;
        P = static_cast<char *>(seqa_cuda_host_alloc(Want));
        if (!P) throw std::runtime_error(std::string("seqa_cuda_host_alloc: ") + seqa_cuda_last_error());
        Cap = Want;
    }
};

// Non-owning array view with the vector accessors the result type needs.
template <typename T> struct Span {
    T *P = nullptr;
    size_t N = 0;
    size_t size() const { return N; }
    bool empty() const { return N == 0; }
    T *data() const { return P; }
    T *begin() const { return P; }
    T *end() const { return P + N; }
    T &operator[](size_t I) const { return P[I]; }
};

// Scores + op strings of a batch without any std::list.  The arrays live in one block of
// page-locked memory shared with (and recycled by) the aligner that produced them: copying a PackedAlignments is
// cheap and keeps the block alive.
struct PackedAlignments {
    Span<int32_t> Score;
    Span<uint32_t> StartI, StartJ, EndI, EndJ, OpsLen;
    Span<uint64_t> OpsOff;
    // 0 diagonal, 1 up (Seq1 symbol vs Blank), 2 left (Blank vs Seq2 symbol); in the 2-bit wire format
    // (SEQA_FLAG_OPS_2BIT, what the aligners below request): 4 ops per byte, OpsOff in bytes, OpsLen in ops
    Span<uint8_t> Ops;
    bool TwoBit = false;
    std::shared_ptr<PinnedBlock> Store; // owns the memory behind the spans
    size_t size() const { return Score.size(); }
    // op K of pair P, whatever the wire format
    unsigned op(size_t P, uint32_t K) const
    {
        return TwoBit ? (Ops[OpsOff[P] + (K >> 2)] >> (2 * (K & 3))) & 3u : Ops[OpsOff[P] + K];
    }
    // CIGAR-like run-length view of pair P: (op, run length) pairs in alignment order
    std::vector<std::pair<uint8_t, uint32_t>> runs(size_t P) const
    {
        std::vector<std::pair<uint8_t, uint32_t>> R;
        if (OpsLen[P] == SEQA_PAIR_UNSUPPORTED) return R;
        for (uint32_t K = 0; K < OpsLen[P]; K++) {
            const uint8_t O = (uint8_t)op(P, K);
            if (!R.empty() && R.back().first == O)
                R.back().second++;
            else
                R.emplace_back(O, 1u);
        }
        return R;
    }
};

namespace detail {

template <typename C> inline const char *bytes_of(const C &S)
{
    static_assert(sizeof(typename C::value_type) == 1, "the GPU path aligns 8-bit symbols");
    return S.size() ? reinterpret_cast<const char *>(&*S.begin()) : "";
}

template <typename Ty, typename Fn> inline bool is_equality_on_bytes(Fn &F)
{
    for (int A = 0; A < 256; A++)
        for (int B = 0; B < 256; B++)
            if (F(static_cast<Ty>(static_cast<unsigned char>(A)), static_cast<Ty>(static_cast<unsigned char>(B))) != (A == B)) return false;
    return true;
}

// Runs Body(Lo, Hi, T) over [0, N) cut into contiguous ranges, one per host thread (the calling thread takes the
// first); small N runs inline.  The header's packing, offset prefix and list expansion all go through this.
inline size_t hostThreads(size_t Work, size_t MinPerThread)
{
    const size_t Hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    return std::max<size_t>(1, std::min<size_t>(std::min<size_t>(Hw, 32), Work / std::max<size_t>(MinPerThread, 1)));
}
template <typename BodyTy> inline void parallelFor(size_t N, size_t Threads, BodyTy Body)
{
    if (Threads <= 1 || N == 0) {
        Body((size_t)0, N, (size_t)0);
        return;
    }
    std::vector<std::thread> Pool;
    Pool.reserve(Threads - 1);
    for (size_t T = 1; T < Threads; T++) Pool.emplace_back(Body, N * T / Threads, N * (T + 1) / Threads, T);
    Body((size_t)0, N / Threads, (size_t)0);
    for (std::thread &Th : Pool) Th.join();
}

#ifdef SEQA_HAVE_AVX2_DISPATCH
// 32 symbols per step where the CPU has AVX2 (checked at run time; the header itself needs no -mavx2): codes by shift + and,
// validity by a byte shuffle through "ACTG", four codes per byte by two multiply-adds (1,4 then 1,16).  Packs the first
// whole sequence (the last partial step from a padded stack copy); *Bad is set when a symbol is not one of ACGT.
__attribute__((target("avx2"))) inline uint64_t pack2bit_step_avx2(const __m256i X, __m256i *Acc) // 32 symbols -> 8 packed bytes
{
    const __m256i Lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i Three = _mm256_set1_epi8(3), W1 = _mm256_set1_epi16(0x0401), W2 = _mm256_set1_epi32(0x00100001);
    const __m256i Pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i C = _mm256_and_si256(_mm256_srli_epi16(X, 1), Three);
    *Acc = _mm256_or_si256(*Acc, _mm256_xor_si256(_mm256_shuffle_epi8(Lut, C), X));
    const __m256i P = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(C, W1), W2), Pick);
    return (uint64_t)(uint32_t)_mm256_extract_epi32(P, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(P, 4) << 32);
}
__attribute__((target("avx2"))) inline size_t pack2bit_avx2(const char *S, size_t Len, uint8_t *Dst, uint64_t *Bad)
{
    __m256i Acc = _mm256_setzero_si256();
    size_t K = 0;
    for (; K + 32 <= Len; K += 32) {
        const uint64_t V = pack2bit_step_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(S + K)), &Acc);
        std::memcpy(Dst + (K >> 2), &V, 8);
    }
    if (K < Len) { // the last 1..31 symbols, padded with 'A' (code 0) in a stack buffer: one more step
        alignas(32) char Tail[32];
        std::memset(Tail, 'A', 32);
        std::memcpy(Tail, S + K, Len - K);
        const uint64_t V = pack2bit_step_avx2(_mm256_load_si256(reinterpret_cast<const __m256i *>(Tail)), &Acc);
        std::memcpy(Dst + (K >> 2), &V, (Len - K + 3) >> 2);
        K = Len;
    }
    if (!_mm256_testz_si256(Acc, Acc)) *Bad |= 1;
    return K;
}
#endif

// 2-bit packing of one sequence for SEQA_FLAG_BASES_2BIT (4 symbols per byte, A0 C1 T2 G3 = (letter >> 1) & 3), eight
// symbols per step (32 with AVX2).  Returns false when a symbol outside ACGT is met (the batch then goes out as 8-bit symbols).
inline bool pack2bit(const char *S, size_t Len, uint8_t *Dst)
{
    uint64_t Bad = 0;
    size_t K = 0;
#ifdef SEQA_HAVE_AVX2_DISPATCH
    static const bool HasAvx2 = __builtin_cpu_supports("avx2");
    if (HasAvx2 && Len >= 8) {
        pack2bit_avx2(S, Len, Dst, &Bad);
        return Bad == 0;
    }
#endif
    for (; K + 8 <= Len; K += 8) {
        uint64_t X;
        std::memcpy(&X, S + K, 8);
        const uint64_t C = (X >> 1) & 0x0303030303030303ull;
        // the letter every code stands for, rebuilt bit by bit, must give back the input
        const uint64_t B0 = C & 0x0101010101010101ull, B1 = (C >> 1) & 0x0101010101010101ull, T = B1 & ~B0;
        Bad |= X ^ 0x4141414141414141ull ^ (T | (B0 << 1) | (B1 << 2) | (T << 4));
        uint64_t Y = (C | (C >> 6)) & 0x000F000F000F000Full; // 16-bit lanes: c0 | c1 << 2
        Y = (Y | (Y >> 12)) & 0x000000FF000000FFull;          // 32-bit lanes: four codes per byte
        const uint16_t Out = (uint16_t)(Y | (Y >> 24));
        std::memcpy(Dst + (K >> 2), &Out, 2);
    }
    for (; K < Len; K += 4) {
        unsigned V = 0;
        for (size_t Q = 0; Q < 4 && K + Q < Len; Q++) {
            const unsigned Ch = (unsigned char)S[K + Q], Code = (Ch >> 1) & 3u;
            Bad |= (uint64_t)(Ch ^ (unsigned char)"ACTG"[Code]);
            V |= Code << (2 * Q);
        }
        Dst[K >> 2] = (uint8_t)V;
    }
    return Bad == 0;
}

// Is the matching functor an EQUIVALENCE on the 8-bit symbols -- F(a,b) == (class[a] == class[b]), with symbols that do not
// even match themselves in a "matches nothing" class?  Then it is table-driven equality (seqa_batch_in.sym_class) and runs
// on the GPU path: case-insensitive comparison, purine / pyrimidine, amino-acid groups, 'N' that matches nothing...
// Checked exhaustively (65,536 + at most 65,536 calls, once per aligner).  Returns 0 (no), 1 (plain equality: no table
// needed) or 2 (Table holds the classes).
template <typename Ty, typename Fn> inline int classify_functor(Fn &F, std::vector<uint8_t> &Table)
{
    auto Sym = [](int V) { return static_cast<Ty>(static_cast<unsigned char>(V)); };
    Table.assign(256, 0);
    int Rep[256], Classes = 0;
    bool Identity = true;
    for (int A = 0; A < 256; A++) {
        if (!F(Sym(A), Sym(A))) {
            Table[A] = (uint8_t)SEQA_CLASS_NEVER;
            Identity = false;
            continue;
        }
        int K = -1;
        for (int C = 0; C < Classes && K < 0; C++)
            if (F(Sym(A), Sym(Rep[C]))) K = C;
        if (K < 0) {
            if (Classes == 254) { // more classes than ids: only plain equality has that many
                K = Classes; // keep counting to tell equality apart below
            } else {
                Rep[Classes] = A;
                K = Classes++;
            }
        } else {
            Identity = false;
        }
        if (K >= 254) {
            if (!Identity) return 0;
            continue;
        }
        Table[A] = (uint8_t)K;
    }
    if (Identity) { // candidates: every symbol alone in its class -> must be == on all pairs
        for (int A = 0; A < 256; A++)
            for (int B = 0; B < 256; B++)
                if (F(Sym(A), Sym(B)) != (A == B)) return 0;
        return 1;
    }
    for (int A = 0; A < 256; A++)
        for (int B = 0; B < 256; B++)
            if (F(Sym(A), Sym(B)) != (Table[A] == Table[B] && Table[A] != (uint8_t)SEQA_CLASS_NEVER)) return 0;
    return 2;
}

template <typename Fn> inline bool is_null_functor(const Fn &) { return false; }
template <typename R, typename... A> inline bool is_null_functor(const std::function<R(A...)> &F) { return !F; }
template <typename R, typename... A> inline bool is_null_functor(R (*F)(A...)) { return F == nullptr; }
inline bool is_null_functor(std::nullptr_t) { return true; }

} // namespace detail
} // namespace seqa

template <typename ContainerType, typename Ty = typename ContainerType::value_type, Ty Blank = Ty(0),
          typename MatchFnTy = std::function<bool(Ty, Ty)>>
class SequenceAligner {
    ScoringSystem Scoring;
    MatchFnTy Match;
    int EqualityChecked = -1;        // -1 unknown, 0 not usable on the GPU path, 1 equality / nullptr, 2 class table
    std::vector<uint8_t> ClassTable; // EqualityChecked == 2: seqa_batch_in.sym_class
    // page-locked staging, reused from call to call: packed inputs (indices, symbols) and the result block, which is
    // recycled only once the caller has dropped every PackedAlignments that still points into it
    std::shared_ptr<seqa::PinnedBlock> IdxBlock, BasesBlock, OutBlock;

  public:
    using EntryType = typename AlignedSequence<Ty, Blank>::Entry;
    using PairType = std::pair<ContainerType, ContainerType>;

    SequenceAligner(ScoringSystem Scoring, MatchFnTy Match = nullptr) : Scoring(Scoring), Match(Match) {}
    virtual ~SequenceAligner() = default;

    ScoringSystem &getScoring() { return Scoring; }
    bool match(Ty V1, Ty V2) { return seqa::detail::is_null_functor(Match) ? V1 == V2 : Match(V1, V2); }
    MatchFnTy getMatchOperation() { return Match; }
    Ty getBlank() { return Blank; }

    // Scores of the most recent getAlignment / getAlignments call (the reference exposes none).
    std::vector<int> LastScores;
    // Pairs of the most recent batch call that the GPU path rejected one by one (ops_len == SEQA_PAIR_UNSUPPORTED: the
    // three LocalGotoh shapes that are undefined behaviour in the reference, include/SALocalGotoh.h:484-488).  Their
    // AlignedSequence is empty; getAlignment (one pair) throws std::invalid_argument instead.
    std::vector<size_t> LastUnsupported;
    // Input wire format of the most recent call: true = 2-bit symbols (every symbol was one of ACGT), false = 8-bit.
    // ForceByteInputs = true keeps the 8-bit form always (A/B measurements).
    bool LastInputsTwoBit = false, ForceByteInputs = false;

    virtual AlignedSequence<Ty, Blank> getAlignment(ContainerType &Seq0, ContainerType &Seq1) = 0;

    // Pads a local alignment out to a global one (reference include/SequenceAlignment.h:156-189):
    // Seq1[0,Idx1) then Seq2[0,Idx2) as gap columns, the local part, Seq1[EndIdx1,..) then Seq2[EndIdx2,..).
    void forceGlobal(ContainerType &Seq1, ContainerType &Seq2, AlignedSequence<Ty, Blank> &Result, int Idx1, int Idx2,
                     int EndIdx1, int EndIdx2)
    {
        std::list<EntryType> Out;
        for (int I = 0; I < Idx1; I++) Out.emplace_back(Seq1[I], Blank, false);
        for (int J = 0; J < Idx2; J++) Out.emplace_back(Blank, Seq2[J], false);
        Out.splice(Out.end(), Result.Data);
        for (size_t I = (size_t)EndIdx1; I < Seq1.size(); I++) Out.emplace_back(Seq1[I], Blank, false);
        for (size_t J = (size_t)EndIdx2; J < Seq2.size(); J++) Out.emplace_back(Blank, Seq2[J], false);
        Result.Data.swap(Out);
    }

  protected:
    // ---- the boundary: everything below funnels into seqa_cuda_align_batch ----
    void requireGpuEligible()
    {
        if (EqualityChecked < 0)
            EqualityChecked = seqa::detail::is_null_functor(Match) ? 1 : seqa::detail::classify_functor<Ty>(Match, ClassTable);
        if (!EqualityChecked)
            throw std::invalid_argument("seqalib_b200: the GPU path compares symbols with == or through a class table; matching "
                                        "functors that are not an equivalence on the 8-bit symbols are reference-only (see "
                                        "include/SequenceAlignment.h)");
    }

    // First(P) / Second(P) return the two sequences of pair P (no pointer vectors are built for a million pairs).
    //
    // The symbols are packed LAZILY, wave by wave, from the library's own pipeline (seqa_cuda_align_batch_lazy): while the
    // GPU works on wave k the host threads pack wave k+1 -- reading a million std::strings out of the heap costs about as
    // much as the GPU work on them (tests/cpp/bench_header.cpp, SEQA_API_TIMING=1).  Lengths and offsets are laid out up
    // front for the 2-bit wire format; if a symbol outside ACGT turns up, the call is repeated with 8-bit symbols.
    template <typename FirstFn, typename SecondFn> seqa::PackedAlignments runBatch(int Algo, size_t N, FirstFn First, SecondFn Second)
    {
        static_assert(sizeof(Ty) == 1, "seqalib_b200: the GPU path aligns 8-bit symbols (char); wider types are reference-only");
        requireGpuEligible();
        seqa::PackedAlignments R;
        R.TwoBit = true;
        if (N == 0) {
            LastScores.clear();
            LastUnsupported.clear();
            return R;
        }
        auto Up = [](size_t X) { return (X + 63) / 64 * 64; };
        const bool Timing = std::getenv("SEQA_API_TIMING") != nullptr; // phase times of this call on stderr
        const auto T0 = std::chrono::steady_clock::now();
        auto Since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(); };
        // ---- lengths, then offsets for the chosen wire format (threaded two-pass prefix sums) ----
        if (!IdxBlock) IdxBlock = std::make_shared<seqa::PinnedBlock>();
        if (!BasesBlock) BasesBlock = std::make_shared<seqa::PinnedBlock>();
        IdxBlock->reserve(Up(8 * N) * 2 + Up(4 * N) * 2);
        uint64_t *Off1 = reinterpret_cast<uint64_t *>(IdxBlock->P);
        uint64_t *Off2 = reinterpret_cast<uint64_t *>(IdxBlock->P + Up(8 * N));
        uint32_t *Len1 = reinterpret_cast<uint32_t *>(IdxBlock->P + 2 * Up(8 * N));
        uint32_t *Len2 = reinterpret_cast<uint32_t *>(IdxBlock->P + 2 * Up(8 * N) + Up(4 * N));
        const size_t Threads = seqa::detail::hostThreads(N, 8192);
        std::vector<uint64_t> PartSyms(Threads + 1, 0), PartBytes(Threads + 1, 0);
        seqa::detail::parallelFor(N, Threads, [&](size_t Lo, size_t Hi, size_t T) {
            uint64_t Syms = 0, Bytes = 0;
            for (size_t P = Lo; P < Hi; P++) {
                const uint32_t A = (uint32_t)First(P).size(), B = (uint32_t)Second(P).size();
                Len1[P] = A;
                Len2[P] = B;
                Syms += (uint64_t)A + B;
                Bytes += (uint64_t)((A + 3) >> 2) + ((B + 3) >> 2);
            }
            PartSyms[T + 1] = Syms;
            PartBytes[T + 1] = Bytes;
        });
        for (size_t T = 0; T < Threads; T++) {
            PartSyms[T + 1] += PartSyms[T];
            PartBytes[T + 1] += PartBytes[T];
        }
        const uint64_t Total = PartSyms[Threads], TotalPacked = PartBytes[Threads];
        BasesBlock->reserve(Total + 64); // large enough for either wire format
        char *Bases = BasesBlock->P;
        auto LayOut = [&](bool TwoBit) { // Off1 / Off2 of every pair: sequences back to back (byte-aligned ones in the 2-bit format)
            seqa::detail::parallelFor(N, Threads, [&](size_t Lo, size_t Hi, size_t T) {
                uint64_t Run = TwoBit ? PartBytes[T] : PartSyms[T];
                for (size_t P = Lo; P < Hi; P++) {
                    const uint64_t B1 = TwoBit ? (Len1[P] + 3) >> 2 : Len1[P], B2 = TwoBit ? (Len2[P] + 3) >> 2 : Len2[P];
                    Off1[P] = Run;
                    Off2[P] = Run + B1;
                    Run += B1 + B2;
                }
            });
        };
        // ---- results: one page-locked block; the previous one is reused once nobody else holds it ----
        const size_t OpsCap = (size_t)(Total / 4 + N + 1);
        if (!OutBlock || OutBlock.use_count() > 1) OutBlock = std::make_shared<seqa::PinnedBlock>();
        OutBlock->reserve(Up(4 * N) * 6 + Up(8 * N) + Up(OpsCap));
        char *O = OutBlock->P;
        R.Score = {reinterpret_cast<int32_t *>(O), N};
        R.StartI = {reinterpret_cast<uint32_t *>(O + Up(4 * N)), N};
        R.StartJ = {reinterpret_cast<uint32_t *>(O + 2 * Up(4 * N)), N};
        R.EndI = {reinterpret_cast<uint32_t *>(O + 3 * Up(4 * N)), N};
        R.EndJ = {reinterpret_cast<uint32_t *>(O + 4 * Up(4 * N)), N};
        R.OpsLen = {reinterpret_cast<uint32_t *>(O + 5 * Up(4 * N)), N};
        R.OpsOff = {reinterpret_cast<uint64_t *>(O + 6 * Up(4 * N)), N};
        R.Ops = {reinterpret_cast<uint8_t *>(O + 6 * Up(4 * N) + Up(8 * N)), OpsCap};
        R.Store = OutBlock;
        seqa_params Prm{};
        Prm.algo = Algo;
        Prm.gap = Scoring.getGapPenalty();
        Prm.gap_open = Scoring.getGapOpenPenalty();
        Prm.gap_extend = Scoring.getGapExtendPenalty();
        Prm.match = Scoring.getMatchProfit();
        Prm.allow_mismatch = Scoring.getAllowMismatch() ? 1 : 0;
        Prm.mismatch = Scoring.getAllowMismatch() ? Scoring.getMismatchPenalty() : 0;
        Prm.device_first = 0;
        Prm.device_count = 0; // every visible device
        // ---- the call: the library asks for the symbols of one wave at a time (from its producer threads) ----
        struct Filler {
            bool TwoBit;
            char *Bases;
            const uint64_t *Off1, *Off2;
            const uint32_t *Len1, *Len2;
            FirstFn *First;
            SecondFn *Second;
            size_t Threads;
            std::atomic<long long> BusyUs{0}; // time spent packing (callbacks of several devices may run concurrently)
            static int fill(void *User, uint64_t Lo, uint64_t Cnt)
            {
                Filler &F = *static_cast<Filler *>(User);
                const auto A = std::chrono::steady_clock::now();
                std::vector<char> Ok(F.Threads, 1);
                // two threads stay free for the library's own producer / consumer, which feed the GPU meanwhile
                const size_t Use = std::max<size_t>(1, std::min<size_t>(F.Threads > 3 ? F.Threads - 2 : F.Threads, (size_t)Cnt / 2048 + 1));
                seqa::detail::parallelFor((size_t)Cnt, Use, [&](size_t L, size_t H, size_t T) {
                    bool Good = true;
                    for (size_t P = (size_t)Lo + L; P < (size_t)Lo + H && Good; P++) {
                        const char *S1 = seqa::detail::bytes_of((*F.First)(P)), *S2 = seqa::detail::bytes_of((*F.Second)(P));
                        if (F.TwoBit) {
                            Good = seqa::detail::pack2bit(S1, F.Len1[P], reinterpret_cast<uint8_t *>(F.Bases) + F.Off1[P]) &&
                                   seqa::detail::pack2bit(S2, F.Len2[P], reinterpret_cast<uint8_t *>(F.Bases) + F.Off2[P]);
                        } else {
                            if (F.Len1[P]) std::memcpy(F.Bases + F.Off1[P], S1, F.Len1[P]);
                            if (F.Len2[P]) std::memcpy(F.Bases + F.Off2[P], S2, F.Len2[P]);
                        }
                    }
                    Ok[T] = Good ? 1 : 0;
                });
                F.BusyUs += (long long)std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - A).count();
                for (char G : Ok)
                    if (!G) return 1; // a symbol outside ACGT: this batch cannot go out 2-bit packed
                return 0;
            }
        };
        bool TwoBitIn = !ForceByteInputs && EqualityChecked != 2; // a class table is applied to 8-bit symbols
        double TLay = 0, TCall = 0, TPackBusy = 0;
        for (;;) {
            LayOut(TwoBitIn);
            TLay = Since();
            Filler F;
            F.TwoBit = TwoBitIn;
            F.Bases = Bases;
            F.Off1 = Off1;
            F.Off2 = Off2;
            F.Len1 = Len1;
            F.Len2 = Len2;
            F.First = &First;
            F.Second = &Second;
            F.Threads = Threads;
            Prm.flags = SEQA_FLAG_OPS_2BIT | (TwoBitIn ? SEQA_FLAG_BASES_2BIT : 0u); // a quarter of the bytes over PCIe, both ways
            seqa_batch_in In{Bases, Off1, Off2, Len1, Len2, (uint64_t)N, TwoBitIn ? TotalPacked : Total,
                             EqualityChecked == 2 ? ClassTable.data() : nullptr};
            seqa_batch_out Out{R.Score.data(), R.StartI.data(), R.StartJ.data(), R.EndI.data(), R.EndJ.data(), R.Ops.data(),
                               R.OpsOff.data(), R.OpsLen.data(), (uint64_t)OpsCap, 0};
            const int Rc = seqa_cuda_align_batch_lazy(&Prm, &In, &Out, &Filler::fill, &F);
            TPackBusy = (double)F.BusyUs.load() / 1e3;
            if (Rc == SEQA_OK) break;
            if (TwoBitIn && Rc == SEQA_ERR_INVALID) { // the filler met a symbol outside ACGT: once more, as 8-bit symbols
                TwoBitIn = false;
                continue;
            }
            throw std::runtime_error(std::string("seqa_cuda_align_batch: ") + seqa_cuda_last_error());
        }
        LastInputsTwoBit = TwoBitIn;
        TCall = Since();
        // (large batches are processed in waves whose op strings sit at each wave's own base offset inside Ops)
        LastScores.resize(N);
        LastUnsupported.clear();
        std::vector<std::vector<size_t>> Rejected(Threads);
        seqa::detail::parallelFor(N, Threads, [&](size_t Lo, size_t Hi, size_t T) {
            std::memcpy(LastScores.data() + Lo, R.Score.data() + Lo, (Hi - Lo) * sizeof(int));
            for (size_t P = Lo; P < Hi; P++)
                if (R.OpsLen[P] == SEQA_PAIR_UNSUPPORTED) Rejected[T].push_back(P);
        });
        for (const std::vector<size_t> &V : Rejected) LastUnsupported.insert(LastUnsupported.end(), V.begin(), V.end());
        if (Timing)
            std::fprintf(stderr, "[seqa api] %zu pairs, %zu threads: lengths + layout %.2f ms, seqa_cuda_align_batch_lazy %.2f ms (packing %s inside: "
                                 "%.2f ms busy), scores %.2f ms\n", N, Threads, TLay, TCall - TLay, TwoBitIn ? "2-bit" : "8-bit", TPackBusy, Since() - TCall);
        return R;
    }

    // op string -> AlignedSequence; Local adds the forceGlobal framing
    AlignedSequence<Ty, Blank> expand(const seqa::PackedAlignments &R, size_t P, ContainerType &Seq1, ContainerType &Seq2, bool Local)
    {
        AlignedSequence<Ty, Blank> Res;
        if (R.OpsLen[P] == SEQA_PAIR_UNSUPPORTED) return Res; // rejected pair: see LastUnsupported
        size_t I = R.StartI[P], J = R.StartJ[P];
        for (uint32_t K = 0; K < R.OpsLen[P]; K++) {
            const unsigned Op = R.op(P, K);
            if (Op == SEQA_OP_DIAG) {
                Res.Data.emplace_back(Seq1[I], Seq2[J], EqualityChecked == 2 ? (bool)Match(Seq1[I], Seq2[J]) : Seq1[I] == Seq2[J]);
                I++, J++;
            } else if (Op == SEQA_OP_UP) {
                Res.Data.emplace_back(Seq1[I], Blank, false);
                I++;
            } else {
                Res.Data.emplace_back(Blank, Seq2[J], false);
                J++;
            }
        }
        if (Local) forceGlobal(Seq1, Seq2, Res, (int)R.StartI[P], (int)R.StartJ[P], (int)R.EndI[P], (int)R.EndJ[P]);
        return Res;
    }

    AlignedSequence<Ty, Blank> alignOne(int Algo, bool Local, ContainerType &Seq1, ContainerType &Seq2)
    {
        seqa::PackedAlignments R = runBatch(Algo, 1, [&](size_t) -> const ContainerType & { return Seq1; },
                                            [&](size_t) -> const ContainerType & { return Seq2; });
        if (!LastUnsupported.empty())
            throw std::invalid_argument("seqalib_b200: this (len1,len2) shape is undefined behaviour in the reference's LocalGotohSA "
                                        "(include/SALocalGotoh.h:484-488) and is not aligned on the GPU path");
        return expand(R, 0, Seq1, Seq2, Local);
    }

    std::vector<AlignedSequence<Ty, Blank>> alignMany(int Algo, bool Local, std::vector<PairType> &Pairs)
    {
        seqa::PackedAlignments R = packMany(Algo, Pairs);
        // std::list materialisation (one heap node per aligned column: the reference's own result type is the cost,
        // SURVEY.md 8 a1): every host thread expands a contiguous range of pairs into the pre-sized vector
        std::vector<AlignedSequence<Ty, Blank>> Out(Pairs.size());
        seqa::detail::parallelFor(Pairs.size(), seqa::detail::hostThreads(Pairs.size(), 256), [&](size_t Lo, size_t Hi, size_t) {
            for (size_t P = Lo; P < Hi; P++) Out[P] = expand(R, P, Pairs[P].first, Pairs[P].second, Local);
        });
        return Out;
    }

    seqa::PackedAlignments packMany(int Algo, std::vector<PairType> &Pairs)
    {
        return runBatch(Algo, Pairs.size(), [&](size_t P) -> const ContainerType & { return Pairs[P].first; },
                        [&](size_t P) -> const ContainerType & { return Pairs[P].second; });
    }
};

// One concrete aligner per reference class: same name, same constructors, same getDefaultScoring().
#define SEQA_DEFINE_ALIGNER(NAME, ALGO, LOCAL, DEFAULT_SCORING)                                                              \
    template <typename ContainerType, typename Ty = typename ContainerType::value_type, Ty Blank = Ty(0),                     \
              typename MatchFnTy = std::function<bool(Ty, Ty)>>                                                               \
    class NAME : public SequenceAligner<ContainerType, Ty, Blank, MatchFnTy> {                                                \
        using BaseType = SequenceAligner<ContainerType, Ty, Blank, MatchFnTy>;                                                \
                                                                                                                              \
      public:                                                                                                                 \
        static ScoringSystem getDefaultScoring() { return DEFAULT_SCORING; }                                                  \
        NAME() : BaseType(getDefaultScoring(), nullptr) {}                                                                    \
        NAME(ScoringSystem Scoring, MatchFnTy Match = nullptr) : BaseType(Scoring, Match) {}                                  \
        AlignedSequence<Ty, Blank> getAlignment(ContainerType &Seq1, ContainerType &Seq2) override                           \
        {                                                                                                                     \
            return this->alignOne(ALGO, LOCAL, Seq1, Seq2);                                                                   \
        }                                                                                                                     \
        /* batched entry point: all pairs in one device batch, split statically over the visible GPUs */                      \
        std::vector<AlignedSequence<Ty, Blank>> getAlignments(std::vector<std::pair<ContainerType, ContainerType>> &Pairs)    \
        {                                                                                                                     \
            return this->alignMany(ALGO, LOCAL, Pairs);                                                                       \
        }                                                                                                                     \
        seqa::PackedAlignments getAlignmentsPacked(std::vector<std::pair<ContainerType, ContainerType>> &Pairs)              \
        {                                                                                                                     \
            return this->packMany(ALGO, Pairs);                                                                               \
        }                                                                                                                     \
    };

// default scorings: reference include/SANeedlemanWunsch.h:244-247, include/SASmithWaterman.h:352,
// include/SAGlobalGotoh.h:441, include/SALocalGotoh.h:509, include/SAHirschberg.h:165-168, include/SAMyersMiller.h:406
SEQA_DEFINE_ALIGNER(NeedlemanWunschSA, SEQA_NW, false, ScoringSystem(-1, 2, -1))
SEQA_DEFINE_ALIGNER(SmithWatermanSA, SEQA_SW, true, ScoringSystem(-1, 1, -1))
// The three affine aligners: the reference's default, ScoringSystem(-1, 2, -1), leaves GapOpen / GapExtend
// uninitialised (it reads indeterminate values: no behaviour to match).  Here a default-constructed affine aligner is
// DEFINED: GapOpen = -1, GapExtend = -1, Match = 2, Mismatch = -1 (the same match / mismatch, a gap of k costs -1 - k).
SEQA_DEFINE_ALIGNER(GlobalGotohSA, SEQA_GLOBAL_GOTOH, false, ScoringSystem(-1, -1, 2, -1, true))
SEQA_DEFINE_ALIGNER(LocalGotohSA, SEQA_LOCAL_GOTOH, true, ScoringSystem(-1, -1, 2, -1, true))
SEQA_DEFINE_ALIGNER(HirschbergSA, SEQA_HIRSCHBERG, false, ScoringSystem(-1, 2, -1))
SEQA_DEFINE_ALIGNER(MyersMillerSA, SEQA_MYERS_MILLER, false, ScoringSystem(-1, -1, 2, -1, true))
#undef SEQA_DEFINE_ALIGNER

// Hooks by which the reference's heuristic aligners (BLAT, MUMmer, LocalGotoh) reach the NW hot path
// (reference include/StaticFuncs.h:12-39): NeedlemanWunsch on whole sequences / on a window, appended to Result.
template <typename ContainerType, typename Ty = typename ContainerType::value_type, Ty Blank = Ty(0),
          typename MatchFnTy = std::function<bool(Ty, Ty)>>
class StaticFuncs {
  public:
    static void useNW(ContainerType &Seq1, ContainerType &Seq2, AlignedSequence<Ty, Blank> &Result, ScoringSystem Scoring, MatchFnTy Match)
    {
        Result.Data.clear();
        bridgeNW(Seq1, Seq2, Result, Scoring, 0, 0, (int)Seq1.size(), (int)Seq2.size(), Match);
    }
    static void bridgeNW(ContainerType &Seq1, ContainerType &Seq2, AlignedSequence<Ty, Blank> &Result, ScoringSystem Scoring, int Idx1,
                         int Idx2, int EndIdx1, int EndIdx2, MatchFnTy Match)
    {
        NeedlemanWunschSA<ArrayView<ContainerType>, Ty, Blank, MatchFnTy> NW(Scoring, Match);
        ArrayView<ContainerType> V1(Seq1), V2(Seq2);
        V1.sliceWindow((size_t)Idx1, (size_t)EndIdx1);
        V2.sliceWindow((size_t)Idx2, (size_t)EndIdx2);
        AlignedSequence<Ty, Blank> Part = NW.getAlignment(V1, V2);
        Result.splice(Part);
    }

    // Batched form of bridgeNW: the anchor-chaining aligners spend their DP time in many small
    // NW calls between anchors (reference include/SAMummer.h:49,75,101, include/SABLAT.h:399).  Collect the windows
    // of one or many alignments and submit them as ONE batch to the GPU path; Results[k] receives window k appended,
    // exactly as bridgeNW(..., Windows[k]) would have.
    struct Window {
        int Idx1, Idx2, EndIdx1, EndIdx2;
    };
    static void bridgeNWBatch(ContainerType &Seq1, ContainerType &Seq2, const std::vector<Window> &Windows,
                              std::vector<AlignedSequence<Ty, Blank>> &Results, ScoringSystem Scoring, MatchFnTy Match)
    {
        typedef ArrayView<ContainerType> View;
        NeedlemanWunschSA<View, Ty, Blank, MatchFnTy> NW(Scoring, Match);
        std::vector<std::pair<View, View>> Pairs;
        Pairs.reserve(Windows.size());
        for (const Window &W : Windows) {
            View V1(Seq1), V2(Seq2);
            V1.sliceWindow((size_t)W.Idx1, (size_t)W.EndIdx1);
            V2.sliceWindow((size_t)W.Idx2, (size_t)W.EndIdx2);
            Pairs.emplace_back(V1, V2);
        }
        std::vector<AlignedSequence<Ty, Blank>> Parts = NW.getAlignments(Pairs);
        if (Results.size() < Windows.size()) Results.resize(Windows.size());
        for (size_t K = 0; K < Windows.size(); K++) Results[K].splice(Parts[K]);
    }
};

#endif // SEQA_SEQUENCE_ALIGNMENT_H
