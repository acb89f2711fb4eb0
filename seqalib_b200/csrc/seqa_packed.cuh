// Inter-sequence fast path for short linear-gap pairs (NeedlemanWunschSA / SmithWatermanSA, reference
// include/SANeedlemanWunsch.h:40-231, include/SASmithWaterman.h:47-339): one THREAD owns TWO pairs, packed
// as the two signed 16-bit halves of every register, and the whole recurrence runs on the sm_100a packed
// integer instructions:
//
//     sim  = PRMT(T0_j, T1_j, sel_i)                 one byte-permute = both pairs' match/mismatch score
//     lg   = VIADD.16x2(Hleft, gap)
//     t    = VIADDMNMX.S16x2[.RELU](Hdiag, sim, lg)   max(Hdiag+sim, Hleft+gap [,0])
//     H    = VIADDMNMX.S16x2(Hup, gap, t)             max(Hup+gap, t)          <- the only serial op per row
//
// NeedlemanWunsch runs COLUMN-NORMALISED, K(i,j) = H(i,j) - j*gap: the left candidate H(i,j-1)+gap becomes K(i,j-1) itself,
// the diagonal one K(i-1,j-1) + (sim - gap) (the profile is biased by -gap), the upper one K(i-1,j) + gap -- the VIADD
// disappears and a cell pair costs PRMT + 2 VIADDMNMX.  (SmithWaterman's clamp at 0 would become a clamp at -j*gap, one
// more instruction: it stays in H-space with the RELU form.)  The trace holds the low bits of K and the walk tests the
// same three equalities, shifted by the same constants; H(M,N) = K(M,N) + N*gap.
//
// T0_j/T1_j are the "column profile" of the two pairs (4 int8 scores, one per possible row base) and sel_i is
// a per-row PRMT selector (row base of pair 0 in nibble 0, of pair 1 in nibble 2, sign-extension nibbles 1/3),
// both precomputed once per batch by pk_prep_kernel, so no per-cell base comparison exists at all.
//
// A thread sweeps its matrix in strips of R rows held in registers; the strip's bottom row crosses to the
// next strip through a private column of shared memory.  Instead of 2-bit direction codes (which would cost
// several compare/select slots per cell) the kernel stores the LOW BYTE of every H value: one PRMT packs
// 2 rows x 2 pairs, one 128-bit store per 8 rows, laid out so a warp writes 512 contiguous bytes.  Adjacent
// cells differ by far less than 128 (checked on the host against the scoring parameters), so the walk kernel
// recovers exact neighbour values from the low bytes and applies the reference's own equality tests and
// priority (diag > up > left) -- bit-exact by construction, no direction derivation on the hot loop.
#pragma once
#include <type_traits>
#include "seqa_common.cuh"

#define PK_NULL 0xffffffffu
#define PK_BND_AHEAD 8 /* column groups between the L2 prefetch of a global strip-boundary piece and its use */
#define PK_BLOCK 128

struct PkWarpJob {
    uint32_t first;   // position of this warp's 64 pairs in perm[]
    uint32_t Mw, Nw;  // max len1 / len2 over the 64 pairs
    uint32_t nstrips; // ceil(Mw / R)
    uint64_t trace_off;  // byte offset of the warp's trace region
    uint64_t prof_off;   // uint2 index of the column profile: uint4 [ceil(Nw/4)][2][32] = {T0,T1} of columns (0,1) / (2,3) per lane
    uint64_t rowsel_off; // uint32 index of the row selectors  [nstrips*R][32], followed by the walk's 2-bit symbol codes (pk_rowsel_elems)
    uint64_t last_off;   // uint4 index of the last-column values (SmithWaterman) [nstrips][R/4][32]
};

// uint32 elements of a job's row-selector region: the selectors [nstrips*R][32], then the 2-bit symbol codes pk_walk2_kernel
// compares instead of the 8-bit symbols -- ROW codes [nstrips][2][32] (16 rows of one pair per word: strip s, pair half k,
// lane) and COLUMN codes [ceil(Ng/4)][2][32] (16 columns per word), both written by pk_prep_kernel (R = 16).
__host__ __device__ inline uint64_t pk_rowsel_elems(uint32_t nstrips, uint32_t Nw, int R)
{
    const uint32_t Ng = (Nw + 3) / 4;
    return (uint64_t)nstrips * (uint64_t)R * 32ull + (uint64_t)nstrips * 64ull + (uint64_t)((Ng + 3) / 4) * 64ull;
}

struct PkArgs {
    const uint8_t *bases;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const uint32_t *perm; // pair ids, PK_NULL = padding
    const PkWarpJob *jobs;
    uint32_t njobs;
    uint2 *prof;
    uint32_t *rowsel;
    uint4 *lastcol; // SW: exact H of every row in the last (padded) column, 4 rows x 2 pairs per uint4
    uint8_t *trace;
    int32_t *score;
    uint32_t *end_i, *end_j, *start_i, *start_j;
    uint8_t *slots;
    const uint64_t *slot_off;
    uint32_t *slot_start, *ops_len;
    int *bad; // set when a base outside ACGT is met
    uint8_t *badpair; // ... and per pair: badpair[p] = 1 (the pair is re-run on the 8-bit kernels, the rest of the batch stays)
    int gap, match, mismatch, allow;
    uint32_t smem_cols; // columns of shared boundary storage per thread
    uint64_t npos;      // njobs * 64
    // affine kernels (seqa_packed_affine.cuh)
    int go, ge;
    int prof_bias;         // subtracted from every profile score (affine: GapOpen + GapExtend)
    uint4 *bound;          // per-warp strip boundary rows (kernels that keep them in global memory)
    uint64_t bound_stride; // uint4 per warp
    uint32_t *ticket;      // job counter: warps draw jobs (largest first) instead of striding over them
    uint32_t prep_stage;   // pk_prep_kernel: bytes of shared staging per warp (multiple of 16; 0 = read global memory directly)
    int colcodes;          // != 0: prep writes 2-bit COLUMN CODES (uint16 [Ng][32] per job, at the job's profile offset) instead of
                           // the column profiles, and the fill builds the profile words itself (kernels that are HBM-bound)
};

// next job of this warp: dynamic (ticket) so that ragged batches, sorted largest-first, balance across warps
__device__ __forceinline__ uint32_t pk_next_job(const PkArgs &A, int lane)
{
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(A.ticket, 1u);
    return __shfl_sync(SEQA_FULL, w, 0);
}

// 2-bit code of an upper-case DNA letter: A0 C1 T2 G3 ((c>>1)&3); valid only for the four letters.
__device__ __forceinline__ unsigned pk_code(unsigned c) { return (c >> 1) & 3u; }
__device__ __forceinline__ bool pk_is_acgt(unsigned c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// ---- prep: column profiles + row selectors ----------------------------------------------------------------
// Sequences are read as ALIGNED 32-bit words (one load per 4 symbols; a funnel shift restores the pair's own
// alignment): a warp's lanes read 32 different sequences, so the number of load instructions, not bytes, is the
// cost.  Words may reach 3 bytes before / 11 bytes after a sequence; `bases` is 4-byte aligned with 16 bytes of slack.
// The loads run PK_PREP_AHEAD words ahead of their use (a ring of registers, indexed at compile time by the unrolled
// loops of pk_prep_kernel): with one word ahead every iteration waited a full HBM round trip (ncu: 45 % of the stall
// samples on the four moves behind the loads, profiles/r02_ncu_pk_prep_fill_walk_sw150_1M.txt); reads reach at most
// 4 * (PK_PREP_AHEAD + 1) bytes behind a sequence (`bases` is allocated with 64 bytes of slack).
#define PK_PREP_AHEAD 4
template <bool SM> // SM: the words come from the warp's TMA-staged copy in shared memory (pk_prep_kernel), else from global memory
struct PkSeqReader {
    const uint32_t *w; // aligned word pointer
    unsigned sh;       // bit shift of the first symbol inside *w
    uint32_t carry, ahead[PK_PREP_AHEAD];
    __device__ __forceinline__ void init(const uint8_t *base, const uint8_t *p)
    {
        const uint64_t o = (uint64_t)(p - base);
        w = reinterpret_cast<const uint32_t *>(base) + (o >> 2);
        sh = (unsigned)(o & 3u) * 8u;
        carry = ld(w++);
#pragma unroll
        for (int k = 0; k < PK_PREP_AHEAD; k++) ahead[k] = ld(w++);
    }
    static __device__ __forceinline__ uint32_t ld(const uint32_t *q) { return SM ? *q : __ldg(q); }
    template <int K> __device__ __forceinline__ uint32_t next4() // the next 4 symbols, first in the low byte; K = call index % PK_PREP_AHEAD
    {
        const uint32_t hi = ahead[K];
        ahead[K] = ld(w++);
        const uint32_t v = sh ? ((carry >> sh) | (hi << (32u - sh))) : carry;
        carry = hi;
        return v;
    }
};

// One job (64 pairs, two per lane).  `base` is the 4-byte aligned start of the memory the four sequences s* live in: `bases`
// itself, or the warp's staged copy of the job's span of it (SM).
template <bool SM>
__device__ __forceinline__ void pk_prep_job(const PkArgs &A, const int R, const PkWarpJob &J, const int lane, const uint32_t p0,
                                            const uint32_t p1, const uint32_t M0, const uint32_t N0, const uint32_t M1, const uint32_t N1,
                                            const uint8_t *base, const uint8_t *sa0, const uint8_t *sb0, const uint8_t *sa1,
                                            const uint8_t *sb1)
{
    const unsigned mm = A.allow ? ((unsigned)(A.mismatch - A.prof_bias) & 0xffu) : 0x80u; // -128 marks "never" (see header)
    const unsigned mt = (unsigned)(A.match - A.prof_bias) & 0xffu;
    const unsigned mm4 = mm * 0x01010101u, mx = mt ^ mm;
    {
        PkSeqReader<SM> a0, b0, a1, b1;
        a0.init(base, sa0);
        b0.init(base, sb0);
        a1.init(base, sa1);
        b1.init(base, sb1);
        // Four symbols per 32-bit word are handled together: codes4 = (w >> 1) & 0x03030303 (A0 C1 T2 G3 in every byte); the
        // word is valid iff "ACTG"[code] gives back every byte (one PRMT table look-up + XOR); bytes behind the end of
        // the sequence are masked out.
        unsigned bad0 = 0, bad1 = 0; // pair p0 / pair p1 met a symbol outside ACGT
        const uint32_t Ng = (J.Nw + 3) >> 2;
        auto valid_mask = [](uint32_t have) { return have >= 4u ? 0xffffffffu : ((1u << (8u * have)) - 1u); }; // have = symbols left
        auto check4 = [&](uint32_t w, uint32_t codes, uint32_t vmask, unsigned &bad) {
            const unsigned x = codes | (codes >> 4);                    // byte 0: c0 | c1 << 4, byte 2: c2 | c3 << 4
            const unsigned letters = seqa_prmt(0x47544341u, 0u, seqa_prmt(x, 0u, 0x4420)); // "ACTG"[c0..c3]
            bad |= (letters ^ w) & vmask;
        };
        uint32_t cacc0 = 0, cacc1 = 0; // 2-bit codes of the 16 columns / rows in flight (pair 0 / pair 1)
        uint32_t *wcode = A.rowsel + J.rowsel_off + (uint64_t)J.nstrips * (uint32_t)R * 32 + lane; // [nstrips][2][32], then [ceil(Ng/4)][2][32]
        uint4 *pout = reinterpret_cast<uint4 *>(A.prof + J.prof_off) + lane; // [cg][half][lane]: every warp store / load is 512 contiguous bytes
        uint16_t *cout = reinterpret_cast<uint16_t *>(A.prof + J.prof_off) + lane; // colcodes: [cg][lane], byte k = pair k, 2 bits per column
        auto colstep = [&](uint32_t cg, auto kk) {
            constexpr int K = decltype(kk)::value;
            const uint32_t j0 = cg * 4;
            const uint32_t h0 = j0 < N0 ? N0 - j0 : 0u, h1 = j0 < N1 ? N1 - j0 : 0u; // real columns left in this group
            const uint32_t w0 = h0 ? b0.template next4<K>() : 0u, w1 = h1 ? b1.template next4<K>() : 0u;
            const uint32_t k0 = (w0 >> 1) & 0x03030303u, k1 = (w1 >> 1) & 0x03030303u;
            check4(w0, k0, valid_mask(h0), bad0);
            check4(w1, k1, valid_mask(h1), bad1);
            // four 2-bit codes per byte: the walk's column codes (16 columns per word), and the fills' with A.colcodes
            uint32_t y0 = (k0 | (k0 >> 6)) & 0x000f000fu, y1 = (k1 | (k1 >> 6)) & 0x000f000fu;
            y0 = (y0 | (y0 >> 12)) & 0xffu;
            y1 = (y1 | (y1 >> 12)) & 0xffu;
            cacc0 |= y0 << (8 * K);
            cacc1 |= y1 << (8 * K);
            if (A.colcodes) { // pk_colprof rebuilds the profile words; padding is masked there
                cout[(uint64_t)cg * 32] = (uint16_t)(y0 | (y1 << 8));
                return;
            }
            const uint32_t s0 = k0 << 3, s1 = k1 << 3; // 8 * code: the byte position of the matching row base in the profile
            unsigned t[8];
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                // padded column: every score -128
                t[2 * c] = c < h0 ? mm4 ^ (mx << ((s0 >> (8 * c)) & 0xffu)) : 0x80808080u;
                t[2 * c + 1] = c < h1 ? mm4 ^ (mx << ((s1 >> (8 * c)) & 0xffu)) : 0x80808080u;
            }
            pout[(uint64_t)cg * 64] = make_uint4(t[0], t[1], t[2], t[3]);
            pout[(uint64_t)cg * 64 + 32] = make_uint4(t[4], t[5], t[6], t[7]);
        };
        for (uint32_t cg = 0; cg < Ng; cg += PK_PREP_AHEAD) { // unrolled by the ring depth: the ring index is a compile-time constant
            colstep(cg, std::integral_constant<int, 0>());
            if (cg + 1 < Ng) colstep(cg + 1, std::integral_constant<int, 1>());
            if (cg + 2 < Ng) colstep(cg + 2, std::integral_constant<int, 2>());
            if (cg + 3 < Ng) colstep(cg + 3, std::integral_constant<int, 3>());
            uint32_t *o = wcode + (uint64_t)J.nstrips * 64 + (uint64_t)(cg >> 2) * 64;
            o[0] = cacc0;
            o[32] = cacc1;
            cacc0 = cacc1 = 0;
        }
        const uint32_t rows = J.nstrips * (uint32_t)R;
        uint32_t *rout = A.rowsel + J.rowsel_off + lane;
        auto rowstep = [&](uint32_t i0, auto kk) {
            constexpr int K = decltype(kk)::value;
            const uint32_t h0 = i0 < M0 ? M0 - i0 : 0u, h1 = i0 < M1 ? M1 - i0 : 0u;
            const uint32_t w0 = h0 ? a0.template next4<K>() : 0u, w1 = h1 ? a1.template next4<K>() : 0u;
            const uint32_t v0 = valid_mask(h0), v1 = valid_mask(h1);
            const uint32_t k0 = (w0 >> 1) & 0x03030303u & v0, k1 = (w1 >> 1) & 0x03030303u & v1; // rows behind the end: code 0
            check4(w0, k0, v0, bad0);
            check4(w1, k1, v1, bad1);
            {
                uint32_t y0 = (k0 | (k0 >> 6)) & 0x000f000fu, y1 = (k1 | (k1 >> 6)) & 0x000f000fu;
                cacc0 |= ((y0 | (y0 >> 12)) & 0xffu) << (8 * K);
                cacc1 |= ((y1 | (y1 >> 12)) & 0xffu) << (8 * K);
            }
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                const unsigned c0 = (k0 >> (8 * c)) & 3u, c1 = (k1 >> (8 * c)) & 3u;
                // nibble0: byte c0 of T0; nibble1: its sign (8 | c0); nibble2: byte 4 + c1 (= T1); nibble3: its sign (12 | c1)
                rout[(uint64_t)(i0 + c) * 32] = 0xC480u + c0 * 0x11u + c1 * 0x1100u;
            }
        };
        for (uint32_t i0 = 0; i0 < rows; i0 += 4 * PK_PREP_AHEAD) {
            rowstep(i0, std::integral_constant<int, 0>());
            if (i0 + 4 < rows) rowstep(i0 + 4, std::integral_constant<int, 1>());
            if (i0 + 8 < rows) rowstep(i0 + 8, std::integral_constant<int, 2>());
            if (i0 + 12 < rows) rowstep(i0 + 12, std::integral_constant<int, 3>());
            uint32_t *o = wcode + (uint64_t)(i0 >> 4) * 64; // rows is a multiple of 16 (R = 16)
            o[0] = cacc0;
            o[32] = cacc1;
            cacc0 = cacc1 = 0;
        }
        if (bad0 | bad1) *A.bad = 1;
        if (bad0 && p0 != PK_NULL) A.badpair[p0] = 1;
        if (bad1 && p1 != PK_NULL) A.badpair[p1] = 1;
    }
}

// A.prep_stage != 0 (dynamic shared memory: per warp [A.prep_stage + 64] bytes of staging, then one mbarrier per warp): a job
// whose 128 sequences lie within A.prep_stage bytes of `bases` (64 pairs of a dense batch do) is brought in by ONE TMA bulk
// copy of that span and read from shared memory; the per-lane 4-byte global loads of the other form -- 32 different sequences
// per warp load, every one a sector of its own -- were the kernel's whole cost (long-scoreboard stalls, ncu round 2).
__global__ void __launch_bounds__(PK_BLOCK) pk_prep_kernel(PkArgs A, int R)
{
    SEQA_DYN_SMEM(unsigned char, stage_raw);
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned char *stage = stage_raw + (size_t)wib * (A.prep_stage + 64u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage_raw + (size_t)wpb * (A.prep_stage + 64u)) + wib;
    const bool can_stage = A.prep_stage != 0 && (reinterpret_cast<uintptr_t>(A.bases) & 15u) == 0;
    unsigned phase = 0;
    if (can_stage) {
        if (lane == 0) seqa_mbar_init(bar);
        __syncwarp();
    }
    for (uint32_t w = gw; w < A.njobs; w += nw) {
        const PkWarpJob J = A.jobs[w];
        const uint32_t p0 = A.perm[J.first + 2 * lane], p1 = A.perm[J.first + 2 * lane + 1];
        const uint32_t M0 = p0 == PK_NULL ? 0u : A.len1[p0], N0 = p0 == PK_NULL ? 0u : A.len2[p0];
        const uint32_t M1 = p1 == PK_NULL ? 0u : A.len1[p1], N1 = p1 == PK_NULL ? 0u : A.len2[p1];
        const uint64_t oa0 = p0 == PK_NULL ? 0ull : A.off1[p0], ob0 = p0 == PK_NULL ? 0ull : A.off2[p0];
        const uint64_t oa1 = p1 == PK_NULL ? 0ull : A.off1[p1], ob1 = p1 == PK_NULL ? 0ull : A.off2[p1];
        if (can_stage) {
            // the span of `bases` the job reads: [lo, hi)
            uint64_t lo = ~0ull, hi = 0;
            if (p0 != PK_NULL) lo = min(oa0, ob0), hi = max(oa0 + M0, ob0 + N0);
            if (p1 != PK_NULL) lo = min(lo, min(oa1, ob1)), hi = max(hi, max(oa1 + M1, ob1 + N1));
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                lo = min(lo, __shfl_xor_sync(SEQA_FULL, lo, d));
                hi = max(hi, __shfl_xor_sync(SEQA_FULL, hi, d));
            }
            lo &= ~15ull;
            if (hi > lo && hi - lo <= (uint64_t)A.prep_stage) { // warp-uniform
                const unsigned bytes = (unsigned)((hi - lo + 15ull) & ~15ull); // <= 15 bytes behind hi: inside the slack of `bases`
                if (lane == 0) seqa_bulk_load(stage, A.bases + lo, bytes, bar);
                seqa_mbar_wait(bar, phase);
                phase ^= 1u;
                pk_prep_job<true>(A, R, J, lane, p0, p1, M0, N0, M1, N1, stage, p0 == PK_NULL ? stage : stage + (oa0 - lo),
                                  p0 == PK_NULL ? stage : stage + (ob0 - lo), p1 == PK_NULL ? stage : stage + (oa1 - lo),
                                  p1 == PK_NULL ? stage : stage + (ob1 - lo));
                __syncwarp(); // every lane is done with the staged copy before the next job's bytes land in it
                continue;
            }
        }
        pk_prep_job<false>(A, R, J, lane, p0, p1, M0, N0, M1, N1, A.bases, A.bases + oa0, A.bases + ob0, A.bases + oa1, A.bases + ob1);
    }
}

// bits of x under mask m inserted into w: ONE LOP3 (w, x, m -> (w & ~m) | (x & m), LUT 0xd8 with inputs a = w, b = x, c = m)
#ifdef SEQA_EMU
static inline unsigned pk_insert(unsigned w, unsigned x, unsigned m) { return (w & ~m) | (x & m); }
#else
__device__ __forceinline__ unsigned pk_insert(unsigned w, unsigned x, unsigned m)
{
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xd8;" : "=r"(r) : "r"(w), "r"(x), "r"(m)); // a&~c | b&c: a=0xf0, b=0xcc, c=0xaa -> (0xf0 & 0x55) | (0xcc & 0xaa) = 0x50 | 0x88
    return r;
}
#endif
__device__ __forceinline__ unsigned pk_dup(int v) { return ((unsigned)v & 0xffffu) * 0x00010001u; }
__device__ __forceinline__ int pk_half(unsigned v, int k) { return (int)(int16_t)(k ? (v >> 16) : (v & 0xffffu)); }

// Column profile words of one 4-column group rebuilt from the 2-bit column codes (PkArgs.colcodes): T[2c + k] = the four
// int8 scores of column c of pair k against row bases A/C/T/G = mm4 ^ (mx << 8*code); padded columns score -128 everywhere.
// ~4 ALU instructions per column and pair, once per 16 rows -- for the kernels that are HBM-bound, where re-reading 8 bytes
// of profile per column, pair and strip costs more than rebuilding it (profiles/r02_ncu_pkg_fill*.txt).
__device__ __forceinline__ void pk_colprof(unsigned codes, int left0, int left1, unsigned mm4, unsigned mx, unsigned *T)
{
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned s0 = ((codes >> (2 * c)) & 3u) * 8u, s1 = ((codes >> (8 + 2 * c)) & 3u) * 8u;
        T[2 * c] = c < left0 ? mm4 ^ (mx << s0) : 0x80808080u;
        T[2 * c + 1] = c < left1 ? mm4 ^ (mx << s1) : 0x80808080u;
    }
}

// ---- fill -------------------------------------------------------------------------------------------------
// Trace: the low TB bits (TB = 8, or 4 when Match + |Mismatch| + 2|Gap| <= 7) of every H value.  Layout of a warp
// job (Ng = ceil(Nw/4) column groups): 16-byte pieces of PR = 16/TB rows x 4 columns x 2 pairs,
//   piece(s, cg, rg, lane)  at  trace_off + (((s*Ng + cg)*(R/PR) + rg)*32 + lane) * 16       (rg = r / PR)
//     TB == 8 (2 rows): byte c*4 + (r%2)*2 + k
//     TB == 4: the two pairs are separated (two more PRMT per 8 words) so that a piece belongs to ONE pair and covers
//              8 rows x 4 columns: piece(s, cg, hs, k, lane) at trace_off + ((((s*Ng + cg)*(R/8) + hs)*2 + k)*32 + lane)*16,
//              word ((r%8)/4)*2 + c/2, byte r%4, nibble c%2.
//              (Measured alternatives, profiles/r01_notes_walk_layout.md: pair-major 128-byte lines cut the walk's HBM
//              reads 10.4 -> 6.2 GB but cost the fill more than the walk gains -- scattered stores +41 %, staging
//              through shared memory costs the third resident CTA.)
// (c = column inside the group, k = pair half), so that every warp store is one contiguous 512-byte run and a
// walk step usually stays inside the piece it already holds (thread-major 128-byte lines were measured 2x slower
// in the fill: 32 lines per store instruction saturate the LSU).
// The column profiles of the next 4-column group are prefetched into registers one group ahead.
// PK_PAIR_PIECES (TB == 4 only): the pieces of column groups 2q and 2q+1 of one pair sit side by side in ONE 32-byte
// sector (piece index ((((s*ceil(Ng/2) + cg/2)*(R/8) + hs)*2 + k)*32 + lane)*2 + cg%2): a diagonal walk then changes
// sector every 8 columns instead of every 4.  The fill's warp stores become 32 half-sectors spread over 1 KB, completed
// by the next column group.
#ifndef PK_PAIR_PIECES
#define PK_PAIR_PIECES 0
#endif
__host__ __device__ inline uint64_t pk_trace_bytes(uint32_t nstrips, uint32_t Nw, int R, int TB)
{
    const uint32_t Ng = (Nw + 3) / 4;
    const uint32_t Nge = (PK_PAIR_PIECES == 1 && TB == 4) ? (Ng + 1) / 2 * 2 : Ng;
    return (uint64_t)nstrips * Nge * (uint64_t)(R * TB / 16) * 32ull * 16ull;
}
// 16-byte piece index of (strip s, column group cg, row band rb of the strip, pair half k, lane) for TB == 4
__host__ __device__ inline uint32_t pk_piece4(uint32_t s, uint32_t Ng, uint32_t cg, uint32_t RG, uint32_t rb, uint32_t k, uint32_t lane)
{
    if (RG == 1) return ((s * Ng + cg) * 2u + k) * 32u + lane; // TB == 2: one 16-row piece per strip and column group, one layout
    if (PK_PAIR_PIECES == 2) return ((s * Ng + cg) * 2u + k) * 32u * RG + lane * RG + rb; // the strip's row bands side by side
    if (PK_PAIR_PIECES) return (((((s * ((Ng + 1) / 2) + (cg >> 1)) * RG + rb) * 2u + k) * 32u + lane) * 2u) + (cg & 1u);
    return (((s * Ng + cg) * RG + rb) * 2u + k) * 32u + lane;
}

#ifdef SEQA_EMU
static inline void pk_prefetch_l2_line(const void *) {}
#else
__device__ __forceinline__ void pk_prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
#ifdef SEQA_EMU
static inline void pk_store_stream(uint4 *p, uint4 v) { *p = v; }
static inline void pk_store_stream(uint2 *p, uint2 v) { *p = v; }
#else
__device__ __forceinline__ void pk_store_stream(uint4 *p, uint4 v) { __stcs(p, v); }
__device__ __forceinline__ void pk_store_stream(uint2 *p, uint2 v) { __stcs(p, v); }
#endif

// GB = false: the strip boundary row lives in shared memory (pairs up to 320 columns); GB = true: in a per-warp
// global (L2-resident) row, read one column group ahead -- any length whose scores fit 16 bits.
template <bool LOCAL, int R, int TB, bool GB, bool CODES = false>
__global__ void __launch_bounds__(PK_BLOCK, 3) pk_fill_kernel(PkArgs A)
{
    static_assert(R % 2 == 0, "R must be even");
    static_assert(TB == 2 || TB == 4 || TB == 8, "trace bits");
    SEQA_DYN_SMEM(unsigned, top);
    constexpr int RP = R / 2;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned gap2 = pk_dup(A.gap);
    const unsigned mmb = A.allow ? ((unsigned)(A.mismatch - A.prof_bias) & 0xffu) : 0x80u; // CODES: as pk_prep_kernel
    const unsigned mm4 = mmb * 0x01010101u, mx = ((unsigned)(A.match - A.prof_bias) & 0xffu) ^ mmb;
    // GB: the strip's bottom row, [cg][lane] x 4 columns x 2 pairs, as 8-BIT DIFFERENCES along the row (PK_BND_DELTA): the row is
    // written in strip s and read in strip s+1, ~100 us and hundreds of MB of trace later, so it makes the round trip through
    // HBM (the resident warps' rows outgrow the L2) -- a third on top of a 2-bit trace.  Neighbouring columns of a row differ by
    // [gap, match - gap] (column-normalised NW: [0, match - 2 gap]), inside int8 for every packed scoring (host check
    // match + |mismatch| + 2|gap| <= 120), so 8 bytes per group instead of 16 carry the row exactly; 14 more ALU instructions
    // per group and strip (128 cells).
    uint2 *__restrict__ bnd = reinterpret_cast<uint2 *>(A.bound + (uint64_t)gw * A.bound_stride) + lane;
    for (;;) {
        const uint32_t w = pk_next_job(A, lane);
        if (w >= A.njobs) break;
        const PkWarpJob J = A.jobs[w];
        const uint32_t p0 = A.perm[J.first + 2 * lane], p1 = A.perm[J.first + 2 * lane + 1];
        const int M0 = p0 == PK_NULL ? 0 : (int)A.len1[p0], N0 = p0 == PK_NULL ? 0 : (int)A.len2[p0];
        const int M1 = p1 == PK_NULL ? 0 : (int)A.len1[p1], N1 = p1 == PK_NULL ? 0 : (int)A.len2[p1];
        const int Ng = ((int)J.Nw + 3) >> 2, Nw = (int)J.Nw;
        const uint4 *__restrict__ prof = reinterpret_cast<const uint4 *>(A.prof + J.prof_off) + lane;
        const uint16_t *__restrict__ ccode = reinterpret_cast<const uint16_t *>(A.prof + J.prof_off) + lane;
        const uint32_t *__restrict__ rowsel = A.rowsel + J.rowsel_off + lane;
        uint8_t *__restrict__ trace = A.trace + J.trace_off + (uint64_t)lane * 16;
        int best0 = 0, best1 = 0, bi0 = 0, bi1 = 0; // SW: running (max, last row holding it)
        int corner0 = 0, corner1 = 0;               // NW: H(M,N)
        // row 0 of the matrix = first strip's upper boundary (SW 0, NW j*gap: include/SANeedlemanWunsch.h:61-62)
        if (!GB)
            for (int jj = 0; jj < Nw; jj++) top[jj * PK_BLOCK + tid] = 0u; // SW: H(0,j) = 0; NW: K(0,j) = j*gap - j*gap = 0
        for (int s = 0; s < (int)J.nstrips; s++) {
            const int i0 = s * R;
            const bool first = s == 0, keep = s + 1 < (int)J.nstrips;
            unsigned H[R], sel[R], rmax[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                sel[r] = rowsel[(uint64_t)(i0 + r) * 32];
                H[r] = LOCAL ? 0u : pk_dup((i0 + r + 1) * A.gap); // column 0 (include/SANeedlemanWunsch.h:59-60)
                rmax[r] = 0u;
            }
            unsigned diag = LOCAL ? 0u : pk_dup(i0 * A.gap);
            uint8_t *__restrict__ tr = trace + (uint64_t)s * ((PK_PAIR_PIECES == 1 && TB == 4) ? (Ng + 1) / 2 * 2 : Ng) * (R * TB / 16 * 512);
            uint4 na = make_uint4(0, 0, 0, 0), nb = na; // profile of the next group: {T0,T1} x 4 columns
            unsigned ncode = 0;                          // CODES: column codes of the next group
            if (CODES) {
                ncode = ccode[0];
            } else {
                na = prof[0];
                nb = prof[32];
            }
            uint2 nu = make_uint2(0, 0);  // GB: boundary of the next group (differences)
            if (GB && !first) nu = bnd[0];
            unsigned uprev = diag;                                       // GB: upper boundary value left of the group (column 0: the corner)
            unsigned bprev = LOCAL ? 0u : pk_dup((i0 + R) * A.gap);      // GB: bottom row value left of the group (column 0 of row i0 + R)
            for (int cg = 0; cg < Ng; cg++) {
                uint4 ca = na, cb = nb;
                const uint2 cu = nu;
                if (CODES) {
                    unsigned T[8];
                    pk_colprof(ncode, N0 - cg * 4, N1 - cg * 4, mm4, mx, T);
                    ca = make_uint4(T[0], T[1], T[2], T[3]);
                    cb = make_uint4(T[4], T[5], T[6], T[7]);
                }
                if (cg + 1 < Ng) {
                    if (CODES) {
                        ncode = ccode[(uint64_t)(cg + 1) * 32];
                    } else {
                        na = prof[(uint64_t)(cg + 1) * 64];
                        nb = prof[(uint64_t)(cg + 1) * 64 + 32];
                    }
                    if (GB && !first) nu = bnd[(uint64_t)(cg + 1) * 32];
                }
                // long pairs: the boundary rows of all resident warps outgrow the L2; pull mine back in well ahead
                if (GB && !first && (lane & 15) == 0 && cg + PK_BND_AHEAD < Ng) pk_prefetch_l2_line(&bnd[(uint64_t)(cg + PK_BND_AHEAD) * 32]);
                unsigned up[4];
                if (GB) {
                    if (first) { // matrix row 0 (SW 0; NW H(0,j) = j*gap, include/SANeedlemanWunsch.h:61-62, i.e. K(0,j) = 0)
#pragma unroll
                        for (int c = 0; c < 4; c++) up[c] = 0u;
                    } else { // byte (2c + k) = difference of column c, pair k: sign-extend into the halves, add up along the row
                        up[0] = __vadd2(uprev, seqa_prmt(cu.x, 0u, 0x9180));
                        up[1] = __vadd2(up[0], seqa_prmt(cu.x, 0u, 0xB3A2));
                        up[2] = __vadd2(up[1], seqa_prmt(cu.y, 0u, 0x9180));
                        up[3] = __vadd2(up[2], seqa_prmt(cu.y, 0u, 0xB3A2));
                        uprev = up[3];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) up[c] = (cg * 4 + c < Nw) ? top[(cg * 4 + c) * PK_BLOCK + tid] : 0u; // padded columns: no boundary
                }
                unsigned bot[4];
                unsigned W[RP][TB == 8 ? 4 : TB == 4 ? 2 : 1];
                // the 4 columns of the group; CAP = this group holds the corner column of one of my global alignments
                // (twice per pair): only that rare variant carries the H(M,N) capture code
                auto cols = [&](auto cap) {
                constexpr bool CAP = decltype(cap)::value;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const unsigned T0 = c == 0 ? ca.x : c == 1 ? ca.z : c == 2 ? cb.x : cb.z;
                    const unsigned T1 = c == 0 ? ca.y : c == 1 ? ca.w : c == 2 ? cb.y : cb.w;
                    unsigned hd = diag, hu = up[c];
                    diag = up[c];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const unsigned sim = seqa_prmt(T0, T1, sel[r]);
                        const unsigned hold = H[r];
                        // SW: max(H(i-1,j-1) + sim, H(i,j-1) + gap, 0); NW (column-normalised): max(K(i-1,j-1) + sim - gap, K(i,j-1))
                        const unsigned t = LOCAL ? __viaddmax_s16x2_relu(hd, sim, __vadd2(hold, gap2)) : __viaddmax_s16x2(hd, sim, hold);
                        const unsigned hn = __viaddmax_s16x2(hu, gap2, t);
                        if (r & 1) {
                            const unsigned w8 = seqa_prmt(H[r - 1], hn, 0x6420); // low bytes: [p0 r-1, p1 r-1, p0 r, p1 r]
                            if (TB == 8)
                                W[r >> 1][c] = w8;
                            else if (TB == 2) { // four columns per byte: column c's two bits inserted at 2c (one LOP3 each)
                                const unsigned M = 0x03030303u << (2 * c);
                                W[r >> 1][0] = c == 0 ? w8 : pk_insert(W[r >> 1][0], w8 << (2 * c), M);
                            } else if ((c & 1) == 0)
                                W[r >> 1][c >> 1] = w8;
                            else // low nibbles of column c-1, high nibbles from column c
                                W[r >> 1][c >> 1] = pk_insert(W[r >> 1][c >> 1], w8 << 4, 0xf0f0f0f0u);
                        }
                        H[r] = hn;
                        hu = hn;
                        hd = hold;
                        if (LOCAL && (c & 1)) rmax[r] = __vimax3_s16x2(rmax[r], hold, hn);
                    }
                    if (GB)
                        bot[c] = hu;
                    else if (cg * 4 + c < Nw)
                        top[(cg * 4 + c) * PK_BLOCK + tid] = hu;
                    if (!LOCAL && CAP) {
                        const int j = cg * 4 + c + 1;
                        if (j == N0 || j == N1) {
#pragma unroll
                            for (int r = 0; r < R; r++) {
                                if (j == N0 && i0 + r + 1 == M0) corner0 = pk_half(H[r], 0) + N0 * A.gap; // H = K + j*gap
                                if (j == N1 && i0 + r + 1 == M1) corner1 = pk_half(H[r], 1) + N1 * A.gap;
                            }
                        }
                    }
                }
                };
                {
                    const bool hit = !LOCAL && (((unsigned)(M0 - 1 - i0) < (unsigned)R && (unsigned)(N0 - 1 - cg * 4) < 4u) ||
                                                ((unsigned)(M1 - 1 - i0) < (unsigned)R && (unsigned)(N1 - 1 - cg * 4) < 4u));
                    if (!LOCAL && __any_sync(SEQA_FULL, hit))
                        cols(std::true_type());
                    else
                        cols(std::false_type());
                }
                if (GB && keep) {
                    const unsigned d0 = __vsub2(bot[0], bprev), d1 = __vsub2(bot[1], bot[0]), d2 = __vsub2(bot[2], bot[1]), d3 = __vsub2(bot[3], bot[2]);
                    bprev = bot[3];
                    bnd[(uint64_t)cg * 32] = make_uint2(seqa_prmt(d0, d1, 0x6420), seqa_prmt(d2, d3, 0x6420));
                }
                if (TB == 8) {
                    uint4 *dst = reinterpret_cast<uint4 *>(tr + (uint64_t)cg * (RP * 32 * 16));
#pragma unroll
                    for (int rp = 0; rp < RP; rp++) pk_store_stream(&dst[rp * 32], make_uint4(W[rp][0], W[rp][1], W[rp][2], W[rp][3]));
                } else if (TB == 2) {
                    // one 16-byte piece per pair and column group: 16 rows x 4 columns x 2 bits, byte r = row r
                    static_assert(TB != 2 || R == 16, "2-bit trace pieces hold 16 rows");
                    uint4 *dst = reinterpret_cast<uint4 *>(tr + (uint64_t)cg * (2 * 32 * 16));
                    pk_store_stream(&dst[0], make_uint4(seqa_prmt(W[0][0], W[1][0], 0x6420), seqa_prmt(W[2][0], W[3][0], 0x6420),
                                                         seqa_prmt(W[4][0], W[5][0], 0x6420), seqa_prmt(W[6][0], W[7][0], 0x6420)));
                    pk_store_stream(&dst[32], make_uint4(seqa_prmt(W[0][0], W[1][0], 0x7531), seqa_prmt(W[2][0], W[3][0], 0x7531),
                                                          seqa_prmt(W[4][0], W[5][0], 0x7531), seqa_prmt(W[6][0], W[7][0], 0x7531)));
                } else {
                    static_assert(TB == 8 || R % 8 == 0, "4-bit trace pieces hold 8 rows");
                    // trace already points at this lane's 16 bytes (lane * 16); PAIR_PIECES: lane * 32 + (cg % 2) * 16
                    uint4 *dst = PK_PAIR_PIECES == 1 ? reinterpret_cast<uint4 *>(tr + (uint64_t)(cg >> 1) * (R / 4 * 32 * 32) + lane * 16 + (cg & 1) * 16)
                                 : PK_PAIR_PIECES == 2 ? reinterpret_cast<uint4 *>(tr + (uint64_t)cg * (R / 4 * 32 * 16) + lane * 16 * (R / 8 - 1))
                                                       : reinterpret_cast<uint4 *>(tr + (uint64_t)cg * (R / 4 * 32 * 16));
                    constexpr int PS = PK_PAIR_PIECES == 1 ? 64 : 32; // uint4 stride between consecutive (hs, k) pieces
#pragma unroll
                    for (int hs = 0; hs < R / 8; hs++) {
                        // W[rp][cc]: rows (2rp, 2rp+1) x pairs x column pair cc  ->  per pair: 4 rows per word
                        const unsigned a0 = W[4 * hs][0], a1 = W[4 * hs + 1][0], a2 = W[4 * hs + 2][0], a3 = W[4 * hs + 3][0];
                        const unsigned b0 = W[4 * hs][1], b1 = W[4 * hs + 1][1], b2 = W[4 * hs + 2][1], b3 = W[4 * hs + 3][1];
                        pk_store_stream(&dst[PK_PAIR_PIECES == 2 ? hs : (hs * 2 + 0) * PS], make_uint4(seqa_prmt(a0, a1, 0x6420), seqa_prmt(b0, b1, 0x6420),
                                                                            seqa_prmt(a2, a3, 0x6420), seqa_prmt(b2, b3, 0x6420)));
                        pk_store_stream(&dst[PK_PAIR_PIECES == 2 ? 32 * (R / 8) + hs : (hs * 2 + 1) * PS], make_uint4(seqa_prmt(a0, a1, 0x7531), seqa_prmt(b0, b1, 0x7531),
                                                                            seqa_prmt(a2, a3, 0x7531), seqa_prmt(b2, b3, 0x7531)));
                    }
                }
            }
            if (LOCAL) {
                // exact H of the strip's rows in the last stored column (4*Ng): the walk chains row MaxRow's low bits
                // leftwards from it and stops at the first (= last, include/SASmithWaterman.h:177) column holding MaxScore
                static_assert(R % 4 == 0, "last-column pieces hold 4 rows");
                uint4 *lc = A.lastcol + J.last_off + (uint64_t)s * (R / 4) * 32 + lane;
#pragma unroll
                for (int q = 0; q < R / 4; q++) pk_store_stream(&lc[q * 32], make_uint4(H[4 * q], H[4 * q + 1], H[4 * q + 2], H[4 * q + 3]));
                // last maximum in row-major order (include/SASmithWaterman.h:177): rows ascending, ">="
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = i0 + r + 1;
                    const int v0 = pk_half(rmax[r], 0), v1 = pk_half(rmax[r], 1);
                    if (i <= M0 && v0 >= best0) { best0 = v0; bi0 = i; }
                    if (i <= M1 && v1 >= best1) { best1 = v1; bi1 = i; }
                }
            }
        }
        if (p0 != PK_NULL) {
            A.score[p0] = LOCAL ? best0 : corner0;
            A.end_i[p0] = LOCAL ? (uint32_t)bi0 : (uint32_t)M0;
            if (!LOCAL) A.end_j[p0] = (uint32_t)N0;
        }
        if (p1 != PK_NULL) {
            A.score[p1] = LOCAL ? best1 : corner1;
            A.end_i[p1] = LOCAL ? (uint32_t)bi1 : (uint32_t)M1;
            if (!LOCAL) A.end_j[p1] = (uint32_t)N1;
        }
    }
}

// ---- walk helpers: fewer memory instructions per traceback step ------------------------------------------------------
// A walk step touches the two sequences and the op slot; every lane has its own addresses, so each load / store
// instruction costs one L1 access per lane.  Symbols are therefore read as aligned 32-bit words kept in a register
// (one load per 4 steps along a sequence) and ops leave as 32-bit words (one store per 4 ops; the at most 3 head and
// 3 tail bytes of a slot, which share words with the neighbouring pairs' slots, go out as single bytes).
struct PkSymCache {
    const uint8_t *base; // A.bases (4-byte aligned, 16 bytes of slack behind the last sequence)
    uint64_t off;        // offset of the sequence
    uint32_t key, word;
    __device__ __forceinline__ void init(const uint8_t *b, uint64_t o)
    {
        base = b;
        off = o;
        key = 0xffffffffu;
        word = 0;
    }
    __device__ __forceinline__ unsigned at(int idx) // symbol idx of the sequence
    {
        const uint64_t a = off + (uint64_t)idx;
        const uint32_t k = (uint32_t)(a >> 2);
        if (k != key) {
            word = *reinterpret_cast<const uint32_t *>(base + ((uint64_t)k << 2));
            key = k;
        }
        return (word >> ((unsigned)(a & 3u) * 8u)) & 0xffu;
    }
};

// The same cache with 16-byte lines kept in shared memory (column `tid` of a [4][TPB] word array per sequence): one global
// load per 16 steps along a sequence instead of one per 4.  The walk's L1 is thrashed by trace lines (1 % hit rate), so
// every symbol word load is an L2 round trip on the serial path of a step.
#ifndef PK_SYM16
#define PK_SYM16 1
#endif
struct PkSymCache16 {
    const uint8_t *base; // A.bases (256-byte aligned allocation, 16 bytes of slack behind the last sequence)
    uint64_t off;
    uint32_t key;
    uint32_t *sm; // sm[w * stride]: word w of the cached 16-byte line
    int stride;
    __device__ __forceinline__ void init(const uint8_t *b, uint64_t o, uint32_t *column, int str)
    {
        base = b;
        off = o;
        key = 0xffffffffu;
        sm = column;
        stride = str;
    }
    // (Also measured: the current word kept in a register on top of the shared line -- one LDS per 4 steps instead of
    // one per step.  Two more live registers under the walk's 40-register cap spill inside the loop; NW 150 bp
    // 6.16 -> 6.50 ms per 1 M pairs, SW 6.64 -> 6.54 against 6.41 for this version.)
    __device__ __forceinline__ unsigned at(int idx)
    {
        const uint64_t a = off + (uint64_t)idx;
        const uint32_t k = (uint32_t)(a >> 4);
        if (k != key) {
            const uint4 v = *reinterpret_cast<const uint4 *>(base + ((uint64_t)k << 4));
            sm[0] = v.x;
            sm[stride] = v.y;
            sm[2 * stride] = v.z;
            sm[3 * stride] = v.w;
            key = k;
        }
        return (sm[(int)((a >> 2) & 3u) * stride] >> ((unsigned)(a & 3u) * 8u)) & 0xffu;
    }
};

struct PkOpWriter {
    uint8_t *slots; // A.slots (4-byte aligned)
    uint64_t pos;   // byte offset of the next op + 1 (ops are written back to front)
    uint64_t atop;  // largest multiple of 4 <= end of the slot: bytes at or above it go out one by one
    uint32_t acc;
    __device__ __forceinline__ void init(uint8_t *s, uint64_t slot_end)
    {
        slots = s;
        pos = slot_end;
        atop = slot_end & ~(uint64_t)3;
        acc = 0;
    }
    __device__ __forceinline__ void put(unsigned op)
    {
        const uint64_t a = --pos;
        if (a >= atop) {
            slots[a] = (uint8_t)op;
        } else {
            acc |= op << ((unsigned)(a & 3u) * 8u);
            if ((a & 3u) == 0) {
                *reinterpret_cast<uint32_t *>(slots + a) = acc;
                acc = 0;
            }
        }
    }
    __device__ __forceinline__ void finish() // the bytes of the last, incomplete word
    {
        if (pos < atop)
            for (uint64_t a = pos; (a & 3u) != 0; a++) slots[a] = (uint8_t)(acc >> ((unsigned)(a & 3u) * 8u));
    }
};

// ---- walk -------------------------------------------------------------------------------------------------
// One thread per pair follows the reference's buildResult (include/SANeedlemanWunsch.h:155-231,
// include/SASmithWaterman.h:220-339) on the stored low bits alone: neighbouring cells differ by less than 2^TB / 2
// (host check), so  H(i,j) == H(i-1,j-1) + sim  <=>  low(i,j) - low(i-1,j-1) - sim == 0 (mod 2^TB)  -- the reference's
// own equality tests, in its own order (diag, then up, else left), with no exact neighbour value ever rebuilt.  The
// exact score is only tracked for SmithWaterman's stop test (H == 0) by subtracting each step's contribution.
// Trace pieces (16 B) are cached in shared memory, 4 per thread, direct-mapped by the parity of the piece's row band
// and column group: the 2 x 2 pieces around a cell never collide, and a lookup is a tag compare plus one LDS instead
// of a register select tree.  Word w of slot s of thread t lives at pcw[s*4+w][t]: bank = t % 32, conflict-free.
#define PK_WALK_TPB 256
#ifndef PK_WALK_MINB
#define PK_WALK_MINB 6 /* resident CTAs per SM the walk is compiled for: 7 (36 registers) spills in the step loop -- measured, see notes */
#endif

template <bool LOCAL, int TB, int R>
__global__ void __launch_bounds__(PK_WALK_TPB, PK_WALK_MINB) pk_walk_kernel(PkArgs A)
{
    __shared__ uint32_t pcw[16][PK_WALK_TPB];
    __shared__ uint32_t tag[4][PK_WALK_TPB];
    const int tid = threadIdx.x;
    const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= A.npos) return;
    const uint32_t p = A.perm[pos];
    if (p == PK_NULL) return;
    const PkWarpJob J = A.jobs[pos >> 6];
    const int lane = (int)((pos & 63) >> 1), half = (int)(pos & 1);
    const int M = (int)A.len1[p], N = (int)A.len2[p];
    const uint32_t Ng = (J.Nw + 3) >> 2;
    constexpr unsigned MASK = TB == 8 ? 0xffu : TB == 4 ? 0xfu : 0x3u;
    // rows per piece: 2 (TB 8: 2 rows x 4 columns x 2 pairs), 8 (TB 4: this pair only) or 16 (TB 2: this pair only)
    constexpr int PRSH = TB == 8 ? 1 : TB == 4 ? 3 : 4;
    constexpr uint32_t RG = (uint32_t)(R >> PRSH);
#if PK_SYM16
    __shared__ uint32_t symw[8][PK_WALK_TPB];
    PkSymCache16 a, b;
    a.init(A.bases, A.off1[p], &symw[0][tid], PK_WALK_TPB);
    b.init(A.bases, A.off2[p], &symw[4][tid], PK_WALK_TPB);
#else
    PkSymCache a, b;
    a.init(A.bases, A.off1[p]);
    b.init(A.bases, A.off2[p]);
#endif
    const int gap = A.gap;
    const uint4 *pieces = reinterpret_cast<const uint4 *>(A.trace + J.trace_off);
#pragma unroll
    for (int q = 0; q < 4; q++) tag[q][tid] = 0xffffffffu;
    // low TB bits of H(ii+1, jj+1) (0-based matrix cell ii, jj); see pk_fill_kernel for the piece layout
    auto fetch = [&](int ii, int jj) -> unsigned {
        const int s = ii / R, r = ii - s * R, cg = jj >> 2, c = jj & 3;
        const uint32_t g = ((uint32_t)s * Ng + (uint32_t)cg) * RG + (uint32_t)(r >> PRSH);
        const uint32_t key = TB == 8 ? g * 32u + (uint32_t)lane
                                     : pk_piece4((uint32_t)s, Ng, (uint32_t)cg, RG, (uint32_t)(r >> PRSH), (uint32_t)half, (uint32_t)lane);
        const int slot = (((ii >> PRSH) & 1) << 1) | (cg & 1);
        if (tag[slot][tid] != key) {
            const uint4 v = pieces[key];
            pcw[slot * 4 + 0][tid] = v.x;
            pcw[slot * 4 + 1][tid] = v.y;
            pcw[slot * 4 + 2][tid] = v.z;
            pcw[slot * 4 + 3][tid] = v.w;
            tag[slot][tid] = key;
        }
        const int w = TB == 8 ? c : TB == 4 ? (((r & 7) >> 2) * 2 + (c >> 1)) : (r >> 2);
        const int sh = TB == 8 ? ((r & 1) * 2 + half) * 8 : TB == 4 ? (r & 3) * 8 + (c & 1) * 4 : (r & 3) * 8 + c * 2;
        return (pcw[slot * 4 + w][tid] >> sh) & MASK;
    };
    auto sext = [&](unsigned d) -> int { // signed difference from its low TB bits (TB 2: the window is [-1, 2], see packed_trace_bits)
        return TB == 8 ? (int)(int8_t)(uint8_t)d : TB == 4 ? ((int)((d & 0xfu) ^ 8u) - 8) : ((int)((d + 1u) & 3u) - 1);
    };
    auto border_low = [&](int i, int j) -> unsigned { // i == 0 or j == 0; NW: K(0,j) = 0, K(i,0) = i*gap
        (void)j;
        return LOCAL ? 0u : (unsigned)(i * gap) & MASK;
    };
    PkOpWriter out;
    out.init(A.slots, A.slot_off[p] + (uint64_t)(M + N));
    int i, j, h;
    if (LOCAL) {
        // MaxCol: the last column of row MaxRow holding MaxScore (include/SASmithWaterman.h:177-182); exact values are
        // chained along the row from H(i,0) = 0 through the differences of neighbouring low bits
        const int best = A.score[p];
        i = (int)A.end_i[p];
        int e = 0, bj = N;
        if (i >= 1) {
            // from the right: e = exact H(i, 4*ng) stored by the fill; one step left subtracts the signed difference of
            // the neighbouring low bits.  The first hit is the last column holding MaxScore; for the near-global
            // alignments of the linear-growth regime it sits within a few pieces of the right edge (two 16-byte loads
            // in flight), and in the worst case the whole row is read once, as a left-to-right scan always would.
            const int ii = i - 1, s = ii / R, r = ii - s * R, ng = (int)Ng;
            const uint32_t g0 = (uint32_t)s * Ng * RG + (uint32_t)(r >> PRSH);
            e = (int)reinterpret_cast<const int16_t *>(A.lastcol + J.last_off + ((uint64_t)s * (R / 4) + (uint64_t)(r >> 2)) * 32 + lane)[(r & 3) * 2 + half];
            auto piece = [&](int cg) -> uint4 {
                const uint32_t g = g0 + (uint32_t)max(cg, 0) * RG;
                return pieces[TB == 8 ? g * 32u + (uint32_t)lane
                                      : pk_piece4((uint32_t)s, Ng, (uint32_t)max(cg, 0), RG, (uint32_t)(r >> PRSH), (uint32_t)half, (uint32_t)lane)];
            };
            auto nibbles = [&](const uint4 &v, unsigned *nib) { // the row's 4 columns of a piece
                if (TB == 8) {
                    const int sh = ((r & 1) * 2 + half) * 8;
                    nib[0] = (v.x >> sh) & 0xffu; nib[1] = (v.y >> sh) & 0xffu;
                    nib[2] = (v.z >> sh) & 0xffu; nib[3] = (v.w >> sh) & 0xffu;
                } else if (TB == 4) {
                    const bool hi = (r & 4) != 0;
                    const int sh = (r & 3) * 8;
                    const unsigned b01 = ((hi ? v.z : v.x) >> sh) & 0xffu, b23 = ((hi ? v.w : v.y) >> sh) & 0xffu;
                    nib[0] = b01 & 0xfu; nib[1] = b01 >> 4; nib[2] = b23 & 0xfu; nib[3] = b23 >> 4;
                } else { // byte r of the piece: four columns x 2 bits
                    const int q = r >> 2;
                    const unsigned b = ((q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w) >> ((r & 3) * 8)) & 0xffu;
                    nib[0] = b & 3u; nib[1] = (b >> 2) & 3u; nib[2] = (b >> 4) & 3u; nib[3] = b >> 6;
                }
            };
            bool found = false;
            uint4 cur = piece(ng - 1), nxt = piece(ng - 2);
            for (int cg = ng - 1; cg >= 0 && !found; cg--) {
                const uint4 nn = piece(cg - 2); // in flight while this piece is examined
                unsigned nib[4], left[4];
                nibbles(cur, nib);
                nibbles(nxt, left);
                const unsigned before = cg > 0 ? left[3] : 0u; // column 0: H(i,0) = 0
#pragma unroll
                for (int c = 3; c >= 0; c--) {
                    const int jj = cg * 4 + c; // 0-based column; e = H(i, jj+1)
                    if (!found && jj < N && e == best) { bj = jj + 1; found = true; }
                    e -= sext(nib[c] - (c > 0 ? nib[c - 1] : before));
                }
                cur = nxt;
                nxt = nn;
            }
        }
        j = bj;
        h = best;
        A.end_j[p] = (uint32_t)j;
    } else {
        i = M;
        j = N;
        h = A.score[p] - N * gap; // NW: the trace holds K = H - j*gap
    }
    const int dbias = LOCAL ? 0 : gap; // NW: K(i,j) - K(i-1,j-1) = sim - gap on a diagonal step
    unsigned nc = (unsigned)h & MASK; // low bits of H(i,j) / K(i,j)
    while (i > 0 && j > 0) {
        if (LOCAL && h == 0) break; // include/SASmithWaterman.h:281-284
        const bool eq = a.at(i - 1) == b.at(j - 1);
        const int sim = (eq ? A.match : A.mismatch) - dbias;
        const unsigned nd = (i == 1 || j == 1) ? border_low(i - 1, j - 1) : fetch(i - 2, j - 2);
        if ((eq || A.allow) && ((nc - nd - (unsigned)sim) & MASK) == 0) { // H == H(i-1,j-1) + sim, include/SANeedlemanWunsch.h:190
            out.put(0);
            i--; j--;
            h -= sim;
            nc = nd;
            continue;
        }
        const unsigned nu = i == 1 ? border_low(0, j) : fetch(i - 2, j - 1);
        if (((nc - nu - (unsigned)gap) & MASK) == 0) { // H == H(i-1,j) + Gap, include/SANeedlemanWunsch.h:216
            out.put(1);
            i--;
            nc = nu;
        } else { // :223
            nc = j == 1 ? border_low(i, 0) : fetch(i - 1, j - 2);
            out.put(2);
            j--;
        }
        h -= gap;
    }
    if (!LOCAL) { // borders: column 0 -> up, row 0 -> left
        while (i > 0) { out.put(1); i--; }
        while (j > 0) { out.put(2); j--; }
    }
    out.finish();
    const int k = (int)(out.pos - A.slot_off[p]);
    A.start_i[p] = (uint32_t)i;
    A.start_j[p] = (uint32_t)j;
    A.slot_start[p] = (uint32_t)k;
    A.ops_len[p] = (uint32_t)(M + N - k);
}
