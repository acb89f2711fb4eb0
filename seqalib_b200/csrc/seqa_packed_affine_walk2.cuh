// pkg_walk2_kernel -- traceback of the packed AFFINE path (4-bit trace planes), round-synchronous like pk_walk2_kernel
// (seqa_packed_walk2.cuh): a warp alternates a LOAD phase (every lane fetches the trace piece it waits for, the G piece its
// diagonal is about to enter and new symbol codes; op words leave) and a STEP phase of PK_WALK2_STEPS iterations that touch
// shared memory and registers only.  One lookup per iteration and lane:
//
//   state 0  H: diagonal test on G(i-1,j-1)      (include/SAGlobalGotoh.h:286)      -> step diagonally, or state 1
//   state 1  H: is H == Ix?  on Ix(i,j)          (:355 before :411)                  -> state 2 (Ix) or 3 (Iy), no step
//   state 2  Ix: emit "up";   extend? on Ix(i-1,j)  (:336 before :344)               -> stay, or back to state 0
//   state 3  Iy: emit "left"; extend? on Iy(i,j-1)  (:394 before :402)               -> stay, or back to state 0
//
// `val` is the EXACT value of the matrix the lane is in (H, Ix or Iy at the current cell), the planes hold low nibbles of
// G = H + go + ge, Ix and Iy; every test is the reference's own, in its order, as in pkg_walk_kernel.
// Piece cache per thread: 8 slots x 16 bytes -- G: row-band parity x column-pair parity (the next column pair is prefetched
// while the current one is walked), Ix and Iy: row-band parity.
#pragma once
#include "seqa_packed_affine.cuh"
#include "seqa_packed_walk2.cuh"

#ifndef PKG_WALK2_STEPS
#define PKG_WALK2_STEPS PK_WALK2_STEPS /* lookups per STEP phase (PkOpWriter64: at most 5) */
#endif

template <bool LOCAL, int R>
__device__ __forceinline__ void pkg_walk2_pairs(const PkArgs &A, const uint64_t pos, const PkSmemCol S)
{
    static_assert(R == 16, "16-row strips");
    constexpr uint32_t ROWB = PK_WALK2_TPB * 4; // bytes per row of the shared array: rows 0-31 piece words (slot*4 + word), 32-39 tags
    constexpr unsigned MASK = 0xfu;
    constexpr uint32_t NONE = 0xffffffffu;
    const uint32_t p = A.perm[pos];
    const bool have = p != PK_NULL;
    const PkWarpJob J = A.jobs[pos >> 6];
    const int lane = (int)((pos & 63) >> 1), half = (int)(pos & 1);
    const int M = have ? (int)A.len1[p] : 0, N = have ? (int)A.len2[p] : 0;
    const uint32_t NC = (J.Nw + 1) >> 1;
    const int go = A.go, ge = A.ge, gogo = go + ge;
    // piece (s, jc, plane, hf) at uint4 index ((((s*NC + jc)*3 + plane)*2 + hf)*32 + lane: both pairs of a fill lane share it
    const uint4 *pieces = reinterpret_cast<const uint4 *>(A.trace + J.trace_off) + lane;
    const uint32_t *rcode = A.rowsel + J.rowsel_off + (uint64_t)J.nstrips * R * 32 + half * 32 + lane; // [strip][2][32]
    const uint32_t *ccode = rcode + (uint64_t)J.nstrips * 64;                                          // [column block of 16][2][32]
    // tag of a piece: plane << 30 | row band (8 rows) << 16 | column pair
    auto piece_of = [&](uint32_t t) -> uint64_t {
        const uint32_t band = (t >> 16) & 0x3fffu, jc = t & 0xffffu, plane = t >> 30;
        return (uint64_t)((((band >> 1) * NC + jc) * 3u + plane) * 2u + (band & 1u)) * 32;
    };
    auto slot_of = [&](uint32_t t) -> uint32_t { // G: 0-3 (band parity, column-pair parity); Ix: 4-5; Iy: 6-7 (band parity)
        const uint32_t plane = t >> 30, b = (t >> 16) & 1u;
        return plane == 0u ? ((b << 1) | (t & 1u)) : (2u + 2u * plane + b);
    };
    auto put_piece = [&](uint32_t t, const uint4 &v) {
        const uint32_t sl = slot_of(t);
        S.st((sl * 4 + 0) * ROWB, v.x);
        S.st((sl * 4 + 1) * ROWB, v.y);
        S.st((sl * 4 + 2) * ROWB, v.z);
        S.st((sl * 4 + 3) * ROWB, v.w);
        S.st((32 + sl) * ROWB, t);
    };
#pragma unroll
    for (int q = 0; q < 8; q++) S.st((32 + q) * ROWB, NONE);
    const uint64_t slot_begin = have ? A.slot_off[p] : 0;
    PkOpWriter64 out;
    out.init(A.slots, slot_begin, (uint32_t)(M + N));
    int i = 0, j = 0, val = 0;
    if (have) {
        if (LOCAL) {
            // MaxCol: the last column of row MaxRow holding MaxScore (include/SALocalGotoh.h:220-225); exact values are chained
            // from H(i,0) = 0 through the low nibbles of G = H + go + ge, four pieces (8 columns) in flight
            const int best = A.score[p];
            i = (int)A.end_i[p];
            int e = 0, bj = N;
            if (i >= 1) {
                const uint32_t ii = (uint32_t)(i - 1), band = ii >> 3;
                const uint64_t base = (uint64_t)(((band >> 1) * NC) * 3u * 2u + (band & 1u)) * 32; // plane 0; + jc * 6 * 32
                const uint32_t w = (ii & 7u) >> 1, sh0 = ((ii & 1u) * 2u + (uint32_t)half) * 8u;
                const uint32_t ncp = ((uint32_t)N + 1u) >> 1;
                for (uint32_t jc0 = 0; jc0 < ncp; jc0 += 4) {
                    uint32_t wd[4];
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++) {
                        const uint4 v = pieces[base + (uint64_t)min(jc0 + q, ncp - 1u) * 192];
                        wd[q] = w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w;
                    }
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++) {
#pragma unroll
                        for (uint32_t c = 0; c < 2; c++) {
                            const int jj = (int)((jc0 + q) * 2 + c) + 1;
                            if (jj <= N) {
                                const unsigned lowg = (wd[q] >> (sh0 + c * 4u)) & MASK;
                                e += (int)(((lowg - (unsigned)(e + gogo)) & MASK) ^ 8u) - 8;
                                if (e == best) bj = jj;
                            }
                        }
                    }
                }
            }
            j = bj;
            val = best;
            A.end_j[p] = (uint32_t)j;
        } else {
            i = M;
            j = N;
            val = A.score[p];
        }
    }
    const bool allow = A.allow != 0;
    // lane state: 0-3 (above); +4 blocked (waits for the LOAD phase); 8 done
    int st = (!have || i <= 0 || j <= 0 || (LOCAL && val == 0)) ? 8 : 0; // include/SALocalGotoh.h:289,334
    int ci = i - 1, cj = j - 1; // the current cell, 0-based (>= 0 while st < 8)
    uint32_t need = NONE;       // blocked: tag of the piece to load
    uint32_t aw = 0, bw = 0, aw2 = 0, bw2 = 0; // 2-bit symbol codes: current and next 16-row strip / 16-column block (pk_walk2_kernel)
    int cf = 0;
    for (;;) {
        // ---- LOAD
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
            // prefetch (H states): the G piece left of the one the next diagonal test reads -- two diagonal steps away
            uint32_t t2 = NONE;
            if ((st & 3) <= 1 && ci >= 2 && cj >= 3) {
                t2 = seqa_prmt(((uint32_t)(cj - 1) >> 1) - 1u, (uint32_t)(ci - 2) >> 3, 0x5410);
                if (t2 == t1 || S.ld((32 + slot_of(t2)) * ROWB) == t2) t2 = NONE;
            }
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
        // ---- STEP
#pragma unroll 1
        for (int k = 0; k < PKG_WALK2_STEPS; k++) {
            if (st < 4) {
                const bool eq = ((pk_shr_wrap(aw, (uint32_t)ci * 2u) ^ pk_shr_wrap(bw, (uint32_t)cj * 2u)) & 3u) == 0u;
                const int t = val - (eq ? A.match : A.mismatch); // state 0: H(i-1,j-1) if this cell came from the diagonal
                const uint32_t plane = st == 0 ? 0u : (st == 3 ? 2u : 1u);
                const int ni = ci - ((st == 0) | (st == 2)), nj = cj - ((st == 0) | (st == 3)); // the cell the lookup reads
                const bool border = (ni | nj) < 0;
                const bool skip = st == 0 && !(eq || allow); // no diagonal candidate at all: straight to state 1
                const uint32_t band = (uint32_t)ni >> 3, jc = (uint32_t)nj >> 1;
                const uint32_t tagv = seqa_prmt(jc, band, 0x5410) | (plane << 30);
                const uint32_t sl = plane == 0u ? (((band & 1u) << 1) | (jc & 1u)) : (2u + 2u * plane + (band & 1u));
                const uint32_t tg = S.ld((32 + sl) * ROWB);
                const uint32_t wv = S.ld((sl * 4 + (((uint32_t)ni & 7u) >> 1)) * ROWB);
                if (!border && !skip && tg != tagv) { // the piece is not held: wait for the LOAD phase, then repeat this test
                    need = tagv;
                    st |= 4;
                } else {
                    const unsigned low = pk_shr_wrap(wv, (((uint32_t)ni & 1u) * 2u + (uint32_t)half) * 8u + ((uint32_t)nj & 1u) * 4u) & MASK;
                    const int expv = st == 0 ? t + gogo : (st == 1 ? val : val - ge);
                    bool res = low == ((unsigned)expv & MASK);
                    if (border) // state 0: H(i-1,j-1) on the border (0 / go + k*ge); states 2, 3: Ix(0,j) / Iy(i,0) = -10000
                        res = st == 0 ? (t == ((LOCAL || (ci | cj) == 0) ? 0 : go + (ci == 0 ? cj : ci) * ge)) : (val == PKG_NEG + ge);
                    if (skip) res = false;
                    const bool take = st >= 2 || (st == 0 && res);
                    const int nval = st == 0 ? t : (res ? val - ge : val - gogo);
                    const int nst = st == 0 ? (res ? 0 : 1) : (st == 1 ? (res ? 2 : 3) : (res ? st : 0));
                    if (take) {
                        out.put(st == 0 ? 0u : (unsigned)(st - 1));
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
                        val = nval;
                    }
                    st = nst;
                    if (take && (border || (LOCAL && nst == 0 && nval == 0))) st = 8;
                }
            }
        }
        if (__all_sync(SEQA_FULL, st == 8)) break;
    }
    i = ci + 1;
    j = cj + 1;
    if (!have) return;
    if (!LOCAL) { // borders: column 0 -> up (include/SAGlobalGotoh.h:312), row 0 -> left (:370)
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

template <bool LOCAL, int R>
__global__ void __launch_bounds__(PK_WALK2_TPB, PK_WALK2_MINB) pkg_walk2_kernel(PkArgs A)
{
    __shared__ uint32_t sm[40][PK_WALK2_TPB];
    const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= A.npos) return; // npos is a multiple of 64: whole warps leave
    PkSmemCol S;
    S.init(&sm[0][threadIdx.x]);
    pkg_walk2_pairs<LOCAL, R>(A, pos, S);
}
