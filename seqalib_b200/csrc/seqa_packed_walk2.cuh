// pk_walk2_kernel -- traceback of the packed linear path (2- and 4-bit traces), ROUND-SYNCHRONOUS.
//
// Same rules as pk_walk_kernel (reference buildResult, include/SANeedlemanWunsch.h:155-231,
// include/SASmithWaterman.h:220-339, evaluated on the stored low bits; one thread per pair), different execution shape.
// pk_walk_kernel loads a trace piece at the point of the step that misses it: a warp then waits a full HBM round trip
// whenever ANY of its lanes misses, up to five times per step (three neighbour lookups, two symbol lines) -- 166 step
// iterations per warp, nearly every one stalled, and a 270-instruction step body executed with 14 of 32 lanes
// (profiles/r02_ncu_pk_walk_hot_sass.txt).  Here a warp alternates two phases:
//
//   LOAD   every lane that is blocked fetches the ONE trace piece it needs (16 bytes -> its private shared-memory slot) and,
//          if its cell moved into another 16-row strip / 16-column block, the 2-bit symbol codes of that strip / block
//          (one 32-bit word each, written by pk_prep_kernel) -- all lanes' loads are in flight together: one wait per round
//   STEP   lanes run neighbour tests out of shared memory and registers only; a lane that needs a piece it does not hold
//          records it and idles until the next LOAD phase.  One test per iteration (state machine: diagonal, then up,
//          then left -- the reference's order), so lanes in different states execute the SAME instructions on different
//          data: no divergent copies of the lookup code.
//
// A round serves one piece per lane: ~60 rounds (waits) per 150 bp pair instead of ~166 x several, and the symbols cost two
// shifts and a LOP3 per step instead of two shared-memory line caches.  Lanes of a warp cross column groups in the same
// rounds, so the 128-byte lines their 16-byte pieces share are fetched from HBM while the neighbours still want them.
#pragma once
#include "seqa_packed.cuh"

#ifndef PK_WALK2
#define PK_WALK2 1 /* 0: every packed linear walk on pk_walk_kernel */
#endif
#if PK_PAIR_PIECES != 0 /* the measured layout variants of seqa_packed.cuh keep pk_walk_kernel */
#undef PK_WALK2
#define PK_WALK2 0
#endif
#ifndef PK_WALK2_TPB
#define PK_WALK2_TPB 128 /* threads per CTA of the round-synchronous walks (pk_walk2_kernel, pkg_walk2_kernel): a CTA ends with its
                            longest path, and 128-thread CTAs free their slots sooner -- headline 4,378 / 4,418 / 4,417 GCUPS with 256,
                            4,401 / 4,443 / 4,440 with 128 (three A/B runs), 64: like 128 */
#endif
#ifndef PK_WALK2_MINB
#define PK_WALK2_MINB (1280 / PK_WALK2_TPB) /* 1,280 threads per SM at 48 registers (5 CTAs of 256, 10 of 128): 1,536 threads (40 registers)
                                                spill inside the STEP loop -- 1.57 vs 1.19 ms per 1 M x 150 bp pairs */
#endif
#ifndef PK_WALK2_STEPS
#define PK_WALK2_STEPS 4 /* neighbour tests per STEP phase: 3 / 4 / 6 / 8 measured 1.34 / 1.28 / 1.31 / 1.52 ms per 1 M x 150 bp pairs */
#endif

// MaxCol of SmithWaterman: the last column of row MaxRow holding MaxScore (include/SASmithWaterman.h:177-182), found from
// the right: e = exact H(i, last stored column) written by the fill; one step left subtracts the signed difference of the
// neighbouring low bits (same scan as pk_walk_kernel).
template <int TB, int R>
__device__ __forceinline__ int pk_maxcol_scan(const PkArgs &A, const PkWarpJob &J, const uint4 *pieces /* + half*32 + lane */, int lane, int half,
                                              int i, int N, int best, uint32_t Ng)
{
    constexpr int PRSH = TB == 4 ? 3 : 4;
    constexpr uint32_t RG = (uint32_t)(R >> PRSH);
    const int ii = i - 1, s = ii / R, r = ii - s * R, ng = (int)Ng;
    int e = (int)reinterpret_cast<const int16_t *>(A.lastcol + J.last_off + ((uint64_t)s * (R / 4) + (uint64_t)(r >> 2)) * 32 + lane)[(r & 3) * 2 + half];
    auto piece = [&](int cg) -> uint4 {
        const uint32_t key = ((uint32_t)s * Ng + (uint32_t)max(cg, 0)) * RG + (uint32_t)(r >> PRSH);
        return pieces[(uint64_t)key * 64];
    };
    auto sext = [&](unsigned d) -> int { return TB == 4 ? ((int)((d & 0xfu) ^ 8u) - 8) : ((int)((d + 1u) & 3u) - 1); };
    auto nibbles = [&](const uint4 &v, unsigned *nib) { // the row's 4 columns of a piece
        if (TB == 4) {
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
    int bj = N;
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
    return bj;
}

// ops leave back to front as 32-bit words.  put() only shifts the op into a 64-bit shift register (2 instructions); the
// stores happen in flush(), once per LOAD phase for all lanes together (at most 4 ops arrive between two flushes, at most 3
// stay pending after one): a walk step never branches into a store.  Positions are 32-bit offsets from the 4-byte aligned
// start of the pair's slot; the first and the last word of a slot share their other bytes with the neighbouring pairs' slots
// and go out as single bytes.
#ifdef SEQA_EMU
static inline uint32_t pk_funnel_l(uint32_t lo, uint32_t hi, uint32_t n) { return n ? (hi << n) | (lo >> (32u - n)) : hi; } // n < 32
static inline uint32_t pk_funnel_r(uint32_t lo, uint32_t hi, uint32_t n) { return n ? (lo >> n) | (hi << (32u - n)) : lo; }
#else
__device__ __forceinline__ uint32_t pk_funnel_l(uint32_t lo, uint32_t hi, uint32_t n) { return __funnelshift_l(lo, hi, n); }
__device__ __forceinline__ uint32_t pk_funnel_r(uint32_t lo, uint32_t hi, uint32_t n) { return __funnelshift_r(lo, hi, n); }
#endif
struct PkOpWriter64 {
    uint8_t *base;       // A.slots + (slot begin rounded down to 4)
    uint32_t rel;        // offset of the last op taken (ops go back to front; starts at end)
    uint32_t srel;       // offset of the last op stored: ops [rel, srel) are pending in (hi:lo), byte k = offset rel + k
    uint32_t lo, hi;
    __device__ __forceinline__ void init(uint8_t *slots, uint64_t begin, uint32_t len)
    {
        const uint64_t a0 = begin & ~(uint64_t)3;
        base = slots + a0;
        rel = srel = (uint32_t)(begin - a0) + len;
        lo = hi = 0;
    }
    __device__ __forceinline__ void put(unsigned op)
    {
        hi = pk_funnel_l(lo, hi, 8);
        lo = lo * 256u + op;
        rel--;
    }
    __device__ __forceinline__ uint32_t pending_byte(uint32_t k) const { return (k < 4u ? lo >> (8u * k) : hi >> (8u * (k - 4u))) & 0xffu; }
    __device__ __forceinline__ void flush()
    {
        if (srel & 3u) { // the slot's top word (shared with the next pair's slot): single bytes, once the word's ops are all taken
            const uint32_t a = srel & ~3u;
            if (rel > a) return;
#pragma unroll 1
            for (uint32_t o = a; o < srel; o++) base[o] = (uint8_t)pending_byte(o - rel);
            srel = a;
        }
#pragma unroll 1
        while (srel - rel >= 4u) {
            const uint32_t a = srel - 4u, k0 = a - rel; // k0 <= 3 in the steady state
            const uint32_t w = k0 < 4u ? pk_funnel_r(lo, hi, 8u * k0) : hi >> (8u * (k0 - 4u));
            *reinterpret_cast<uint32_t *>(base + a) = w;
            srel = a;
        }
    }
    __device__ __forceinline__ void finish() // everything still pending, byte by byte (the slot's lowest word is shared too)
    {
        flush();
#pragma unroll 1
        for (uint32_t o = rel; o < srel; o++) base[o] = (uint8_t)pending_byte(o - rel);
        srel = rel;
    }
};

#ifdef SEQA_EMU
static inline uint32_t pk_shr_wrap(uint32_t x, uint32_t n) { return x >> (n & 31u); }
#else
__device__ __forceinline__ uint32_t pk_shr_wrap(uint32_t x, uint32_t n) { return __funnelshift_r(x, 0u, n); } // shift amount mod 32
#endif

struct PkSmemCol {
#ifdef SEQA_EMU
    uint32_t *b;
    void init(uint32_t *col) { b = col; }
    uint32_t ld(uint32_t byte_off) const { return b[byte_off / 4]; }
    void st(uint32_t byte_off, uint32_t v) const { b[byte_off / 4] = v; }
#else
    uint32_t a;
    __device__ __forceinline__ void init(uint32_t *col) { a = (uint32_t)__cvta_generic_to_shared(col); }
    __device__ __forceinline__ uint32_t ld(uint32_t byte_off) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + byte_off) : "memory");
        return v;
    }
    __device__ __forceinline__ void st(uint32_t byte_off, uint32_t v) const
    {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + byte_off), "r"(v) : "memory");
    }
#endif
};

template <bool LOCAL, int TB, int R>
__global__ void __launch_bounds__(PK_WALK2_TPB, PK_WALK2_MINB) pk_walk2_kernel(PkArgs A)
{
    static_assert(R == 16 && (TB == 2 || TB == 4), "pk_walk2_kernel: 16-row strips, one-pair pieces");
    static_assert(PK_WALK2_STEPS <= 5, "PkOpWriter64 holds 8 ops: 3 pending + the new ones");
    // rows 0-15: 4 piece slots x 4 words, row = [cg parity, row-band parity, word(2 bits)]; row 16 + r: the tag of row r's slot
    // (each slot's tag four times: a lookup reads word and tag at one computed address)
    __shared__ uint32_t sm[32][PK_WALK2_TPB];
    constexpr uint32_t ROWB = PK_WALK2_TPB * 4; // bytes per row
    const int tid = threadIdx.x;
    const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= A.npos) return; // npos is a multiple of 64: whole warps leave
    const uint32_t p = A.perm[pos];
    const bool have = p != PK_NULL;
    const PkWarpJob J = A.jobs[pos >> 6];
    const int lane = (int)((pos & 63) >> 1), half = (int)(pos & 1);
    const int M = have ? (int)A.len1[p] : 0, N = have ? (int)A.len2[p] : 0;
    const uint32_t Ng = (J.Nw + 3) >> 2;
    constexpr unsigned MASK = TB == 4 ? 0xfu : 0x3u;
    constexpr int PRSH = TB == 4 ? 3 : 4; // rows per piece: 8 (TB 4) or 16 (TB 2)
    constexpr uint32_t RG = (uint32_t)(R >> PRSH);
    const int gap = A.gap;
    // piece key (s*Ng + cg)*RG + rb  ->  16-byte piece index key*64 + half*32 + lane (pk_piece4, default layout)
    const uint4 *pieces = reinterpret_cast<const uint4 *>(A.trace + J.trace_off) + half * 32 + lane;
    const uint32_t *rcode = A.rowsel + J.rowsel_off + (uint64_t)J.nstrips * R * 32 + half * 32 + lane; // [strip][2][32]
    const uint32_t *ccode = rcode + (uint64_t)J.nstrips * 64;                                          // [column block of 16][2][32]
    PkSmemCol S;
    S.init(&sm[0][tid]);
#pragma unroll
    for (int q = 0; q < 16; q++) S.st((16 + q) * ROWB, 0xffffffffu);
    const uint64_t slot_begin = have ? A.slot_off[p] : 0;
    PkOpWriter64 out;
    out.init(A.slots, slot_begin, (uint32_t)(M + N));
    int i = 0, j = 0, h = 0;
    if (have) {
        if (LOCAL) {
            const int best = A.score[p];
            i = (int)A.end_i[p];
            j = i >= 1 ? pk_maxcol_scan<TB, R>(A, J, pieces, lane, half, i, N, best, Ng) : N;
            h = best;
            A.end_j[p] = (uint32_t)j;
        } else {
            i = M;
            j = N;
            h = A.score[p] - N * gap; // NW: the trace holds K = H - j*gap
        }
    }
    const int dbias = LOCAL ? 0 : gap;  // NW: K(i,j) - K(i-1,j-1) = sim - gap on a diagonal step
    const int dmatch = A.match - dbias, dmis = A.mismatch - dbias;
    const bool allow = A.allow != 0;
    unsigned nc = (unsigned)h & MASK;   // low bits of H(i,j) / K(i,j)
    // lane state: 0 diagonal test, 1 up test, 2 left (the reference's order); +4 blocked (waits for the LOAD phase); 8 done
    int st = (!have || i <= 0 || j <= 0 || (LOCAL && h == 0)) ? 8 : 0; // include/SASmithWaterman.h:281-284
    int ci = i - 1, cj = j - 1;  // the current cell, 0-based (>= 0 while st < 8)
    uint32_t need = 0xffffffffu; // blocked: tag (row band << 16 | column group) of the piece to load; 0xffffffff: symbol codes only
    // 2-bit symbol codes of the current cell's 16-row strip / 16-column block (aw, bw) and of the next ones towards the origin
    // (aw2, bw2: loaded one LOAD phase after a crossing, long before the path can cross again); cf: which of them are loaded
    uint32_t aw = 0, bw = 0, aw2 = 0, bw2 = 0;
    int cf = 0; // bit 0 aw, 1 aw2, 2 bw, 3 bw2
    constexpr int steps = PK_WALK2_STEPS; // neighbour tests per STEP phase (PkOpWriter64 holds what they can emit)
    constexpr uint32_t NONE = 0xffffffffu;
    auto piece_of = [&](uint32_t tagv) -> uint64_t { // 16-byte piece index of a tag (relative to `pieces`)
        const uint32_t band = tagv >> 16, cg = tagv & 0xffffu;
        return (uint64_t)(((band >> (4 - PRSH)) * Ng + cg) * RG + (band & (RG - 1u))) * 64;
    };
    auto slot_row = [&](uint32_t tagv) -> uint32_t { return ((tagv & 1u) << 3) | ((tagv >> 14) & 4u); }; // first row of the tag's slot
    auto put_piece = [&](uint32_t tagv, const uint4 &v) {
        const uint32_t srow = slot_row(tagv);
        S.st((srow + 0) * ROWB, v.x);
        S.st((srow + 1) * ROWB, v.y);
        S.st((srow + 2) * ROWB, v.z);
        S.st((srow + 3) * ROWB, v.w);
#pragma unroll
        for (int q = 0; q < 4; q++) S.st((16 + srow + q) * ROWB, tagv);
    };
    for (;;) {
        // ---- LOAD: the piece a blocked lane waits for, the piece its path is about to enter, and the symbol codes of a new
        //      strip / column block -- all lanes' loads in flight together
        if (st < 8) {
            const uint32_t sa = (uint32_t)ci >> 4, sb = (uint32_t)cj >> 4;
            uint32_t na = aw, nb = bw, na2 = aw2, nb2 = bw2;
            if (cf != 15) {
                if (!(cf & 1)) na = rcode[(uint64_t)sa * 64];
                if (!(cf & 2) && sa > 0u) na2 = rcode[(uint64_t)(sa - 1u) * 64];
                if (!(cf & 4)) nb = ccode[(uint64_t)sb * 64];
                if (!(cf & 8) && sb > 0u) nb2 = ccode[(uint64_t)(sb - 1u) * 64];
            }
            const uint32_t t1 = st >= 4 ? need : NONE;
            // prefetch: the left neighbour of the current cell's piece when a diagonal from the cell leaves the piece through
            // its left edge (row inside the piece > column inside the group); skipped when the slot already holds it
            const uint32_t band = (uint32_t)ci >> PRSH, cg = (uint32_t)cj >> 2;
            uint32_t t2 = seqa_prmt(cg - 1u, band, 0x5410);
            if (cg == 0u || ((uint32_t)ci & ((1u << PRSH) - 1u)) <= ((uint32_t)cj & 3u) || t2 == t1 ||
                S.ld((16 + slot_row(t2)) * ROWB) == t2)
                t2 = NONE;
            uint4 v1 = make_uint4(0, 0, 0, 0), v2 = v1;
            if (t1 != NONE) v1 = pieces[piece_of(t1)];
            if (t2 != NONE) v2 = pieces[piece_of(t2)];
            aw = na;
            bw = nb;
            aw2 = na2;
            bw2 = nb2;
            cf = 15;
            out.flush();
            if (t1 != NONE) put_piece(t1, v1);
            if (t2 != NONE) put_piece(t2, v2);
            st &= 3;
        }
        // ---- STEP: one neighbour test per iteration and lane, shared memory and registers only
#pragma unroll 1
        for (int k = 0; k < steps; k++) {
            if (st < 4) {
                const int dd = 3 - st;                             // 3 diagonal, 2 up, 1 left: bit 1 = row - 1, bit 0 = column - 1
                const int ni = ci - (dd >> 1), nj = cj - (dd & 1); // the neighbour under test (0-based; -1: border)
                const uint32_t row = TB == 4 ? ((((uint32_t)ni >> 1) & 6u) + (((uint32_t)nj >> 1) & 1u) + (((uint32_t)nj & 4u) << 1))
                                             : ((((uint32_t)ni >> 2) & 7u) + (((uint32_t)nj & 4u) << 1));
                const uint32_t sh = TB == 4 ? ((uint32_t)ni * 8u + ((uint32_t)nj & 1u) * 4u) : ((uint32_t)ni * 8u + ((uint32_t)nj & 3u) * 2u); // mod 32
                const uint32_t tagv = seqa_prmt((uint32_t)nj >> 2, (uint32_t)ni >> PRSH, 0x5410); // row band << 16 | column group
                const uint32_t wv = S.ld(row * ROWB);
                const uint32_t tg = S.ld(row * ROWB + 16 * ROWB);
                const bool border = (ni | nj) < 0; // H(0,j) / H(i,0): SW 0; NW K(0,j) = 0, K(i,0) = i*gap
                const unsigned v = border ? (LOCAL ? 0u : (unsigned)((ni + 1) * gap) & MASK) : (pk_shr_wrap(wv, sh) & MASK);
                if (!border && tg != tagv) { // the piece is not held: wait for the LOAD phase, then repeat this test
                    need = tagv;
                    st |= 4;
                } else {
                    const bool eq = ((pk_shr_wrap(aw, (uint32_t)ci * 2u) ^ pk_shr_wrap(bw, (uint32_t)cj * 2u)) & 3u) == 0u;
                    const int expect = st == 0 ? (eq ? dmatch : dmis) : gap;
                    bool ok = ((nc - v - (unsigned)expect) & MASK) == 0u; // H == H(neighbour) + score, include/SANeedlemanWunsch.h:190,216
                    if (st == 0) ok = ok && (eq || allow);
                    if (st == 2) ok = true; // :223
                    if (ok) {
                        out.put((unsigned)st);
                        // the cell leaves the strip / column block of aw / bw: the next words move up, LOAD refills them
                        if ((((uint32_t)ni ^ (uint32_t)ci) & ~15u) != 0u) {
                            aw = aw2;
                            cf &= ~2;
                        }
                        if ((((uint32_t)nj ^ (uint32_t)cj) & ~15u) != 0u) {
                            bw = bw2;
                            cf &= ~8;
                        }
                        ci = ni;
                        cj = nj;
                        nc = v;
                        h -= expect;
                        st = (border || (LOCAL && h == 0)) ? 8 : 0;
                    } else {
                        st++;
                    }
                }
            }
        }
        if (__all_sync(SEQA_FULL, st == 8)) break;
    }
    i = ci + 1;
    j = cj + 1;
    if (!have) return;
    if (!LOCAL) { // borders: column 0 -> up, row 0 -> left
        out.flush();
        while (i > 0) { out.put(1); i--; if ((i & 3) == 0) out.flush(); }
        while (j > 0) { out.put(2); j--; if ((j & 3) == 0) out.flush(); }
    }
    out.finish();
    const uint32_t k = out.rel - (uint32_t)(slot_begin & 3u);
    A.start_i[p] = (uint32_t)i;
    A.start_j[p] = (uint32_t)j;
    A.slot_start[p] = k;
    A.ops_len[p] = (uint32_t)(M + N) - k;
}
