// Inter-sequence fast path for short AFFINE-gap pairs (GlobalGotohSA / LocalGotohSA, reference
// include/SAGlobalGotoh.h:53-422, include/SALocalGotoh.h:56-490): like seqa_packed.cuh one THREAD owns TWO pairs in
// the two signed 16-bit halves of every register.  With G = H + (GapOpen+GapExtend) kept instead of H, a cell is
//
//     sim' = PRMT(T0_j, T1_j, sel_i)                       match/mismatch score MINUS (GapOpen+GapExtend), both pairs
//     ix   = VIADDMNMX.S16x2(Ix_up, GapExtend, G_up)        max(Ix(i-1,j)+ge, H(i-1,j)+go+ge)   include/SAGlobalGotoh.h:172-176
//     iy   = VIADDMNMX.S16x2(Iy_left, GapExtend, G_left)    max(Iy(i,j-1)+ge, H(i,j-1)+go+ge)   :179-183
//     m    = VIMNMX.S16x2(ix, iy)
//     H    = VIADDMNMX.S16x2[.RELU](G_diag, sim', m)        max(H(i-1,j-1)+sim, ix, iy [,0])     :192 / SALocalGotoh.h:216
//     G    = VIADD.16x2(H, go+ge)
//
// = 6 ALU-pipe instructions per two cells, no per-cell comparison.  The strip's bottom row (G and Ix) crosses to
// the next strip through a per-warp boundary row in global memory (L2-resident; read one column group ahead), so
// the pair length is not limited by shared memory.
//
// Traceback: the LOW BYTES of G, Ix and Iy (three planes, 1.5 PRMT per two cells).  The walk keeps the exact
// value of the matrix it is in and evaluates the reference's own equality tests in its own order (diag > Ix > Iy,
// extend before open, include/SAGlobalGotoh.h:260-419) on low bytes: every tested difference is far inside
// (-128, 128) for the scoring ranges the host admits, so equality of low bytes is equality.
#pragma once
#include "seqa_packed.cuh"

#define PKG_NEG (-10000) /* the reference's literal "-infinity" */

__host__ __device__ inline uint64_t pkg_trace_bytes(uint32_t nstrips, uint32_t Nw, int R, int TB)
{
    return (uint64_t)nstrips * (TB == 8 ? Nw : (Nw + 1) / 2) * 3ull * (uint64_t)(R / 8) * 512ull;
}

// Trace piece (16 B = 8 rows x 1 column x 2 pairs of one plane; TB == 4: 8 rows x 2 columns, low nibbles), per warp job:
//   piece(s, jc, plane, hf, lane) at trace_off + ((((s*NC + jc)*3 + plane)*(R/8) + hf)*32 + lane) * 16
//   jc = j-1 and NC = Nw (TB 8) or jc = (j-1)/2 and NC = ceil(Nw/2), nibble (j-1)%2 (TB 4)
//   word (r%8)/2, byte (r%2)*2 + k      (r = row inside the strip, hf = r/8, k = pair half)
// TB == 4 halves the HBM write stream that bounds the 8-bit variant; the host admits it when every difference the
// walk tests stays below 16 (packed_affine_trace_bits in seqa_cuda.cu).
template <bool LOCAL, int R, int TB, bool CODES = true>
__global__ void __launch_bounds__(PK_BLOCK, 3) pkg_fill_kernel(PkArgs A)
{
    static_assert(TB == 4 || TB == 8, "trace bits");
    static_assert(R % 8 == 0, "strips are cut into 8-row trace pieces");
    constexpr int RP = R / 2, RH = R / 8;
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int gogo = A.go + A.ge;
    const unsigned ge2 = pk_dup(A.ge), gogo2 = pk_dup(gogo), neg2 = pk_dup(PKG_NEG);
    const unsigned mmb = A.allow ? ((unsigned)(A.mismatch - A.prof_bias) & 0xffu) : 0x80u; // CODES: as pk_prep_kernel
    const unsigned mm4 = mmb * 0x01010101u, mx = ((unsigned)(A.match - A.prof_bias) & 0xffu) ^ mmb;
    // The strip's bottom row (G and Ix of 4 columns x 2 pairs per group and lane) crosses to the next strip through global
    // memory; the rows of the resident warps (116 MB at 250 bp) outgrow the L2 and make the round trip through HBM, a quarter
    // of this HBM-bound kernel's traffic as 16-bit values.  They travel as 8-BIT fields instead, one uint4 per group and lane:
    //   x, y: G(j) - G(j-1) along the row, int8 (inside [go+ge, match - go - ge]: packed_scoring_ok, m + g <= 120)
    //   z, w: H - Ix, uint8: Ix(i,j) >= H(i-1,j) + go + ge and H(i,j) - H(i-1,j) <= match - go - ge, so 0 <= H - Ix <= m + 2g <= 170
    uint4 *__restrict__ bnd = A.bound + (uint64_t)gw * A.bound_stride + lane; // [cg][lane]
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
        uint4 *__restrict__ trace = reinterpret_cast<uint4 *>(A.trace + J.trace_off) + lane;
        int best0 = 0, best1 = 0, bi0 = 0, bi1 = 0; // local: running (max, last row holding it)
        int corner0 = 0, corner1 = 0;               // global: H(M,N)
        const int nstrips = (int)J.nstrips;
        for (int s = 0; s < nstrips; s++) {
            const int i0 = s * R;
            const bool first = s == 0, keep = s + 1 < nstrips;
            unsigned G[R], Y[R], sel[R], rmax[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                sel[r] = rowsel[(uint64_t)(i0 + r) * 32];
                // column 0: H(i,0) = 0 (local, include/SALocalGotoh.h:77-82) or go + i*ge (global, include/SAGlobalGotoh.h:75-80)
                G[r] = pk_dup((LOCAL ? 0 : A.go + (i0 + r + 1) * A.ge) + gogo);
                Y[r] = neg2;
                rmax[r] = gogo2;
            }
            unsigned diag = pk_dup((LOCAL || i0 == 0 ? 0 : A.go + i0 * A.ge) + gogo); // G(i0, 0)
            const int NC = TB == 8 ? Nw : (Nw + 1) >> 1;
            uint4 *__restrict__ tr = trace + (uint64_t)s * NC * (3 * RH * 32);
            uint4 na = make_uint4(0, 0, 0, 0), nb = na;
            unsigned ncode = 0; // CODES: column codes of the next group (the profile words are rebuilt here, seqa_packed.cuh)
            if (CODES) {
                ncode = ccode[0];
            } else {
                na = prof[0];
                nb = prof[32];
            }
            uint4 nu = make_uint4(0, 0, 0, 0);
            if (!first) nu = bnd[0];
            unsigned gprev = diag;                                                                  // G(i0, 0): the upper row left of the group
            unsigned bprev = pk_dup((LOCAL ? 0 : A.go + (i0 + R) * A.ge) + gogo);                   // G(i0 + R, 0): the bottom row left of the group
            for (int cg = 0; cg < Ng; cg++) {
                uint4 ca = na, cb = nb;
                const uint4 cu = nu;
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
                    if (!first) nu = bnd[(uint64_t)(cg + 1) * 32];
                }
                if (!first && (lane & 7) == 0 && cg + PK_BND_AHEAD < Ng) pk_prefetch_l2_line(&bnd[(uint64_t)(cg + PK_BND_AHEAD) * 32]);
                unsigned upG[4] = {0, 0, 0, 0}, upX[4] = {0, 0, 0, 0}; // G and Ix of the row above the strip
                if (!first) {
                    upG[0] = __vadd2(gprev, seqa_prmt(cu.x, 0u, 0x9180));
                    upG[1] = __vadd2(upG[0], seqa_prmt(cu.x, 0u, 0xB3A2));
                    upG[2] = __vadd2(upG[1], seqa_prmt(cu.y, 0u, 0x9180));
                    upG[3] = __vadd2(upG[2], seqa_prmt(cu.y, 0u, 0xB3A2));
                    gprev = upG[3];
                    upX[0] = __vsub2(__vsub2(upG[0], gogo2), seqa_prmt(cu.z, 0u, 0x4140)); // Ix = H - (H - Ix)
                    upX[1] = __vsub2(__vsub2(upG[1], gogo2), seqa_prmt(cu.z, 0u, 0x4342));
                    upX[2] = __vsub2(__vsub2(upG[2], gogo2), seqa_prmt(cu.w, 0u, 0x4140));
                    upX[3] = __vsub2(__vsub2(upG[3], gogo2), seqa_prmt(cu.w, 0u, 0x4342));
                }
                unsigned outG[4], outX[4];
                unsigned Wg[RP], Wx[RP], Wy[RP];
                // CAP = this group holds the corner column of one of my global alignments: only that rare variant carries
                // the H(M,N) capture code (seqa_packed.cuh)
                auto cols = [&](auto cap) {
                constexpr bool CAP = decltype(cap)::value;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int j = cg * 4 + c + 1;
                    const unsigned T0 = c == 0 ? ca.x : c == 1 ? ca.z : c == 2 ? cb.x : cb.z;
                    const unsigned T1 = c == 0 ? ca.y : c == 1 ? ca.w : c == 2 ? cb.y : cb.w;
                    unsigned gu, xu; // G and Ix of the row above the strip in this column
                    if (first) {     // matrix row 0: H(0,j) = 0 / go + j*ge, Ix(0,j) = -10000
                        gu = pk_dup((LOCAL ? 0 : A.go + j * A.ge) + gogo);
                        xu = neg2;
                    } else {
                        gu = upG[c];
                        xu = upX[c];
                    }
                    unsigned gd = diag;
                    diag = gu;
                    unsigned pg = 0, px = 0, py = 0;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const unsigned sim = seqa_prmt(T0, T1, sel[r]);
                        const unsigned gl = G[r];
                        const unsigned ix = __viaddmax_s16x2(xu, ge2, gu);
                        const unsigned iy = __viaddmax_s16x2(Y[r], ge2, gl);
                        const unsigned m = __vmaxs2(ix, iy);
                        const unsigned hn = LOCAL ? __viaddmax_s16x2_relu(gd, sim, m) : __viaddmax_s16x2(gd, sim, m);
                        const unsigned gn = __vadd2(hn, gogo2);
                        if (r & 1) {
                            const unsigned wg = seqa_prmt(pg, gn, 0x6420); // low bytes: [p0 r-1, p1 r-1, p0 r, p1 r]
                            const unsigned wx = seqa_prmt(px, ix, 0x6420);
                            const unsigned wy = seqa_prmt(py, iy, 0x6420);
                            if (TB == 8 || (c & 1) == 0) {
                                Wg[r >> 1] = wg;
                                Wx[r >> 1] = wx;
                                Wy[r >> 1] = wy;
                            } else { // low nibbles of the even column, high nibbles from the odd one
                                Wg[r >> 1] = (Wg[r >> 1] & 0x0f0f0f0fu) | ((wg << 4) & 0xf0f0f0f0u);
                                Wx[r >> 1] = (Wx[r >> 1] & 0x0f0f0f0fu) | ((wx << 4) & 0xf0f0f0f0u);
                                Wy[r >> 1] = (Wy[r >> 1] & 0x0f0f0f0fu) | ((wy << 4) & 0xf0f0f0f0u);
                            }
                        }
                        pg = gn;
                        px = ix;
                        py = iy;
                        if (LOCAL && (c & 1)) rmax[r] = __vimax3_s16x2(rmax[r], gl, gn);
                        G[r] = gn;
                        Y[r] = iy;
                        gd = gl;
                        gu = gn;
                        xu = ix;
                    }
                    outG[c] = gu;
                    outX[c] = xu;
                    if (TB == 8 ? (j <= Nw) : ((c & 1) && j - 1 <= Nw)) {
                        uint4 *dst = tr + (uint64_t)(TB == 8 ? j - 1 : (j - 1) >> 1) * (3 * RH * 32);
#pragma unroll
                        for (int hf = 0; hf < RH; hf++) {
                            pk_store_stream(&dst[(0 * RH + hf) * 32], make_uint4(Wg[hf * 4], Wg[hf * 4 + 1], Wg[hf * 4 + 2], Wg[hf * 4 + 3]));
                            pk_store_stream(&dst[(1 * RH + hf) * 32], make_uint4(Wx[hf * 4], Wx[hf * 4 + 1], Wx[hf * 4 + 2], Wx[hf * 4 + 3]));
                            pk_store_stream(&dst[(2 * RH + hf) * 32], make_uint4(Wy[hf * 4], Wy[hf * 4 + 1], Wy[hf * 4 + 2], Wy[hf * 4 + 3]));
                        }
                    }
                    if (!LOCAL && CAP) {
                        if (j == N0 || j == N1) {
#pragma unroll
                            for (int r = 0; r < R; r++) {
                                if (j == N0 && i0 + r + 1 == M0) corner0 = pk_half(G[r], 0) - gogo;
                                if (j == N1 && i0 + r + 1 == M1) corner1 = pk_half(G[r], 1) - gogo;
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
                if (keep) {
                    unsigned d[4], e[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        d[c] = __vsub2(outG[c], c ? outG[c - 1] : bprev);
                        e[c] = __vsub2(__vsub2(outG[c], gogo2), outX[c]);
                    }
                    bprev = outG[3];
                    bnd[(uint64_t)cg * 32] = make_uint4(seqa_prmt(d[0], d[1], 0x6420), seqa_prmt(d[2], d[3], 0x6420),
                                                        seqa_prmt(e[0], e[1], 0x6420), seqa_prmt(e[2], e[3], 0x6420));
                }
            }
            if (LOCAL) {
                // last maximum in row-major order (include/SALocalGotoh.h:220-225): rows ascending, ">="
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = i0 + r + 1;
                    const int v0 = pk_half(rmax[r], 0) - gogo, v1 = pk_half(rmax[r], 1) - gogo;
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

// ---- walk -------------------------------------------------------------------------------------------------
// One thread per pair; state machine = reference buildResult (include/SAGlobalGotoh.h:235-422,
// include/SALocalGotoh.h:275-473).  h / x / y are the EXACT values of H / Ix / Iy at the current cell.
template <bool LOCAL, int R, int TB>
__global__ void __launch_bounds__(PK_WALK_TPB) pkg_walk_kernel(PkArgs A)
{
    constexpr unsigned MASK = TB == 8 ? 0xffu : 0xfu;
    // trace pieces cached in shared memory like pk_walk_kernel: per plane two slots, direct-mapped by the parity of
    // the piece's row band (a step touches at most two bands of a plane); word w of slot q at pcw[q*4+w][thread]
    __shared__ uint32_t pcw[24][PK_WALK_TPB];
    __shared__ uint32_t tag[6][PK_WALK_TPB];
    const int tid = threadIdx.x;
    const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= A.npos) return;
    const uint32_t p = A.perm[pos];
    if (p == PK_NULL) return;
    const PkWarpJob J = A.jobs[pos >> 6];
    const int lane = (int)((pos & 63) >> 1), half = (int)(pos & 1);
    const int M = (int)A.len1[p], N = (int)A.len2[p];
    const uint32_t NC = TB == 8 ? J.Nw : (J.Nw + 1) >> 1, RH = (uint32_t)(R / 8);
#if PK_SYM16
    __shared__ uint32_t symw[8][PK_WALK_TPB]; // 16-byte symbol lines of both sequences (PkSymCache16)
    PkSymCache16 a, b;
    a.init(A.bases, A.off1[p], &symw[0][tid], PK_WALK_TPB);
    b.init(A.bases, A.off2[p], &symw[4][tid], PK_WALK_TPB);
#else
    PkSymCache a, b;
    a.init(A.bases, A.off1[p]);
    b.init(A.bases, A.off2[p]);
#endif
    const int go = A.go, ge = A.ge, gogo = go + ge;
    const uint4 *pieces = reinterpret_cast<const uint4 *>(A.trace + J.trace_off);
#pragma unroll
    for (int q = 0; q < 6; q++) tag[q][tid] = 0xffffffffu;
    // low TB bits of plane `plane` (0 G = H+go+ge, 1 Ix, 2 Iy) at matrix cell (i, j), i >= 1, j >= 1
    auto low = [&](int plane, int i, int j) -> unsigned {
        const int ii = i - 1, s = ii / R, r = ii - s * R;
        const uint32_t jc = TB == 8 ? (uint32_t)(j - 1) : (uint32_t)(j - 1) >> 1;
        const uint32_t key = ((((uint32_t)s * NC + jc) * 3u + (uint32_t)plane) * RH + (uint32_t)(r >> 3)) * 32u + (uint32_t)lane;
        const int slot = plane * 2 + ((ii >> 3) & 1);
        if (tag[slot][tid] != key) {
            const uint4 v = pieces[key];
            pcw[slot * 4 + 0][tid] = v.x;
            pcw[slot * 4 + 1][tid] = v.y;
            pcw[slot * 4 + 2][tid] = v.z;
            pcw[slot * 4 + 3][tid] = v.w;
            tag[slot][tid] = key;
        }
        const int sh = ((r & 1) * 2 + half) * 8 + (TB == 4 ? ((j - 1) & 1) * 4 : 0);
        return (pcw[slot * 4 + ((r & 7) >> 1)][tid] >> sh) & MASK;
    };
    auto lowG = [&](int i, int j) -> unsigned { return low(0, i, j); };
    auto lowX = [&](int i, int j) -> unsigned { return low(1, i, j); };
    auto lowY = [&](int i, int j) -> unsigned { return low(2, i, j); };
    auto sext = [&](unsigned d) -> int { // signed difference from its low TB bits
        return TB == 8 ? (int)(int8_t)(uint8_t)d : ((int)((d & 0xfu) ^ 8u) - 8);
    };
    auto borderH = [&](int i, int j) -> int { // i == 0 or j == 0
        if (LOCAL || (i == 0 && j == 0)) return 0;
        return go + (i == 0 ? j : i) * ge;
    };
    PkOpWriter out;
    out.init(A.slots, A.slot_off[p] + (uint64_t)(M + N));
    int i, j, h, x = 0, y = 0, state = 0;
    if (LOCAL) {
        // MaxCol: the last column of row MaxRow holding MaxScore (include/SALocalGotoh.h:220-225); exact values are
        // chained from H(i,0) = 0 through the low bytes of G = H + go + ge
        const int best = A.score[p];
        i = (int)A.end_i[p];
        int e = 0, bj = N;
        if (i >= 1) {
            for (int jj = 1; jj <= N; jj++) {
                e += sext(lowG(i, jj) - (unsigned)(e + gogo));
                if (e == best) bj = jj;
            }
        }
        j = bj;
        h = best;
        A.end_j[p] = (uint32_t)j;
    } else {
        i = M;
        j = N;
        h = A.score[p];
    }
    for (;;) {
        if (LOCAL) {
            if (i <= 0 || j <= 0) break; // include/SALocalGotoh.h:289
        } else {
            if (i == 0 && j == 0) break;
            if (j == 0) { out.put(1); i--; continue; } // include/SAGlobalGotoh.h:312
            if (i == 0) { out.put(2); j--; continue; } // :370
        }
        if (state == 0) {
            if (LOCAL && h == 0) break; // H == max(D,0) == 0 (include/SALocalGotoh.h:334)
            const bool eq = a.at(i - 1) == b.at(j - 1);
            if (eq || A.allow) {
                const int t = h - (eq ? A.match : A.mismatch); // H(i-1,j-1) if this cell came from the diagonal
                const bool isd = (i == 1 || j == 1) ? (t == borderH(i - 1, j - 1))
                                                    : (lowG(i - 1, j - 1) == ((unsigned)(t + gogo) & MASK));
                if (isd) { // include/SAGlobalGotoh.h:286
                    out.put(0);
                    i--; j--;
                    h = t;
                    continue;
                }
            }
            if (lowX(i, j) == ((unsigned)h & MASK)) { // H == Ix (:355), before H == Iy (:411)
                state = 1;
                x = h;
            } else {
                state = 2;
                y = h;
            }
        }
        if (state == 1) {
            out.put(1);
            const bool ext = (i == 1) ? (x == PKG_NEG + ge) : (lowX(i - 1, j) == ((unsigned)(x - ge) & MASK)); // :336 before :344
            if (ext) {
                x -= ge;
            } else {
                h = x - gogo;
                state = 0;
            }
            i--;
        } else {
            out.put(2);
            const bool ext = (j == 1) ? (y == PKG_NEG + ge) : (lowY(i, j - 1) == ((unsigned)(y - ge) & MASK)); // :394 before :402
            if (ext) {
                y -= ge;
            } else {
                h = y - gogo;
                state = 0;
            }
            j--;
        }
    }
    out.finish();
    const int k = (int)(out.pos - A.slot_off[p]);
    A.start_i[p] = (uint32_t)i;
    A.start_j[p] = (uint32_t)j;
    A.slot_start[p] = (uint32_t)k;
    A.ops_len[p] = (uint32_t)(M + N - k);
}
