// Packed variant of the linear-space sweeps (seqa_linspace.cuh): the FORWARD and the REVERSE sweep of one node run
// as the two signed 16-bit halves of one warp wavefront.  They have the same column count and row counts that
// differ by at most one, so a task is "row block rb of both sweeps", and a cell pair costs what the inter-sequence
// kernels pay (seqa_packed.cuh / seqa_packed_affine.cuh):
//
//   Hirschberg  PRMT (both sims from the two column profiles), 2 x VIADDMNMX.S16x2                    = 3 per 2 cells
//               (column-normalised like the packed NeedlemanWunsch fill: K(i,j) = H(i,j) - j*gap makes the left
//               candidate K(i,j-1) itself; the boundary rows in global memory stay absolute H, converted when a
//               chunk is staged / flushed)
//   MyersMiller PRMT, 3 x VIADDMNMX.S16x2, VIMNMX.S16x2, VIADD.16x2 (G = H + go + ge kept, not H)     = 6 per 2 cells
//
// against 5 / 8 int32 instructions per ONE cell.  100 kbp scores do not fit 16 bits, so every lane keeps its values
// relative to a private 32-bit base per half and re-bases at every 32-column chunk; the value handed to the next
// lane is corrected by the difference of the two lanes' bases (exchanged once per chunk), the boundary rows between
// row blocks stay absolute int32 in global memory (converted when staged / flushed).  Within a chunk a lane's values
// move by at most (32 + R + 2) * max|score| -- far inside 16 bits for the scoring range the host admits.
// Requires symbols in {A,C,G,T} (checked once per batch by ls_check_acgt_kernel; otherwise the int32 sweeps run).
#pragma once
#include "seqa_linspace.cuh"
#include "seqa_packed.cuh"

#define LS2_SMEM_INTS (4 * 32 + 128 + 8) /* per warp: inH inX outH outX [32], profiles {T0,T1}[64], lane-31 bases[8] */

__global__ void ls_check_acgt_kernel(const uint8_t *__restrict__ bases, uint64_t n, int *bad)
{
    bool b = false;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x)
        b |= !pk_is_acgt(bases[k]);
    if (b) *bad = 1;
}

__device__ __forceinline__ unsigned ls2_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
__device__ __forceinline__ int ls2_lo(unsigned v) { return (int)(int16_t)(v & 0xffffu); }
__device__ __forceinline__ int ls2_hi(unsigned v) { return (int)(int16_t)(v >> 16); }

// Row block rb of the forward sweep (rows af[0..mf), borders bf) and of the reverse sweep (rows ar[0..mr) taken
// backwards, columns backwards, borders br) of one node.  See ls_block for the chunk / staging / flush protocol.
template <bool AFFINE, int R, bool PARTIAL>
__device__ __forceinline__ void ls_block2(const DevScoring &sc, const Borders &bf, const Borders &br,
                                          const uint8_t *__restrict__ af, int mf, const uint8_t *__restrict__ ar, int mr,
                                          const uint8_t *__restrict__ b, int n, int rb,
                                          const int *inHf, const int *inHr, const int *inXf, const int *inXr, const int *in_prog,
                                          int *outHf, int *outHr, int *outXf, int *outXr, int *out_prog, int *sm)
{
    const int lane = threadIdx.x & 31;
    const int row0 = rb * (32 * R) + lane * R;
    const int nactf = PARTIAL ? min(max(mf - row0, 0), R) : R, nactr = PARTIAL ? min(max(mr - row0, 0), R) : R;
    const int gogo = AFFINE ? sc.go + sc.ge : 0; // affine: registers hold G = H + go + ge
    const unsigned gap2 = pk_dup(sc.gap), ge2 = pk_dup(sc.ge), gogo2 = pk_dup(gogo);
    const int pbias = AFFINE ? gogo : sc.gap; // profile bias: affine G = H + go + ge; linear K = H - j*gap
    const int cnorm = AFFINE ? 0 : sc.gap;    // column normalisation of the linear sweeps
    const unsigned mmb = sc.allow ? ((unsigned)(sc.mismatch - pbias) & 0xffu) : 0x80u; // -128: never (seqa_packed.cuh)
    const unsigned mm4 = mmb * 0x01010101u, mx = ((unsigned)(sc.match - pbias) & 0xffu) ^ mmb;
    unsigned h[R], f[R], sel[R], amask[R];
    int basef = bf.hcolA + (row0 + 1) * bf.hcolB, baser = br.hcolA + (row0 + 1) * br.hcolB;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = row0 + r + 1;
        const unsigned cf = r < nactf ? pk_code(af[i - 1]) : 0u, cr = r < nactr ? pk_code(ar[mr - i]) : 0u;
        sel[r] = cf | ((8u | cf) << 4) | ((4u | cr) << 8) | ((12u | cr) << 12);
        amask[r] = (r < nactf ? 0xffffu : 0u) | (r < nactr ? 0xffff0000u : 0u);
        h[r] = ls2_pack(bf.hcolA + i * bf.hcolB + gogo - basef, br.hcolA + i * br.hcolB + gogo - baser);
        f[r] = AFFINE ? ls2_pack(bf.iyA + i * bf.iyB - basef, br.iyA + i * br.iyB - baser) : 0u;
    }
    unsigned diag_top = ls2_pack(border_hcol(bf, row0) + gogo - basef, border_hcol(br, row0) + gogo - baser);
    unsigned send_h = 0, send_x = 0, adj = 0;
    int *smInH = sm, *smInX = sm + 32, *smOutH = sm + 64, *smOutX = sm + 96;
    uint2 *smT = reinterpret_cast<uint2 *>(sm + 128);
    int *sm31 = sm + 256; // lane 31's bases: [0,1] this chunk (f, r), [2,3] previous chunk
    const uint2 *myT = smT + 32 - lane;
    const int nsteps = n + 31;
    int nHf = 0, nHr = 0, nXf = 0, nXr = 0, known = 0;
    auto wait_for = [&](int need) {
        if (known < need) {
            if (lane == 0) {
                int v;
                while ((v = ls_ld_volatile(in_prog)) < need) ls_pause();
                known = v;
                __threadfence();
            }
            known = __shfl_sync(SEQA_FULL, known, 0);
        }
    };
    auto flush = [&](int j31) { // columns ((j31-1) & ~31) + 1 .. j31 leave as absolute int32 values
        __syncwarp();
        const int base = (j31 - 1) & ~31;
        if (lane < j31 - base) {
            const int c = base + 1 + lane;
            const bool prev = ((c + 30) >> 5) != ((j31 + 30) >> 5); // lane 31 finished this column in the previous chunk
            const int b31f = sm31[prev ? 2 : 0], b31r = sm31[prev ? 3 : 1];
            const unsigned v = (unsigned)smOutH[lane];
            outHf[c] = b31f + ls2_lo(v) - gogo + cnorm * c; // back to absolute H
            outHr[c] = b31r + ls2_hi(v) - gogo + cnorm * c;
            if (AFFINE) {
                const unsigned x = (unsigned)smOutX[lane];
                outXf[c] = b31f + ls2_lo(x);
                outXr[c] = b31r + ls2_hi(x);
            }
        }
        if (out_prog) {
            __threadfence();
            __syncwarp();
            if (lane == 0) ls_st_volatile(out_prog, j31);
        } else {
            __syncwarp();
        }
    };
    auto step = [&](auto chk, int s, int t0) {
        constexpr bool CHECK = decltype(chk)::value;
        unsigned up_h = __vadd2(__shfl_up_sync(SEQA_FULL, send_h, 1), adj);
        unsigned up_x = AFFINE ? __vadd2(__shfl_up_sync(SEQA_FULL, send_x, 1), adj) : 0u;
        const unsigned bh = (unsigned)smInH[s], bx = AFFINE ? (unsigned)smInX[s] : 0u;
        if (lane == 0) {
            up_h = bh;
            up_x = bx;
        }
        const int j = t0 + s - lane + 1;
        if (!CHECK || (j >= 1 && j <= n)) {
            const uint2 T = myT[s];
            unsigned dg = diag_top, uh = up_h, ux = up_x;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const unsigned left = h[r];
                const unsigned sim = seqa_prmt(T.x, T.y, sel[r]);
                unsigned hv, ix = ux;
                if (!AFFINE) {
                    const unsigned tl = __viaddmax_s16x2(dg, sim, left); // max(D, L): K(i-1,j-1) + sim - gap, K(i,j-1)
                    hv = __viaddmax_s16x2(uh, gap2, tl);                                 // max(U, .)
                } else { // registers hold G = H + go + ge: seqa_packed_affine.cuh
                    ix = __viaddmax_s16x2(ux, ge2, uh);
                    const unsigned iy = __viaddmax_s16x2(f[r], ge2, left);
                    hv = __vadd2(__viaddmax_s16x2(dg, sim, __vmaxs2(ix, iy)), gogo2);
                    f[r] = iy;
                }
                if (PARTIAL) { // rows below a sweep hand its last row down, per half
                    hv = (hv & amask[r]) | (uh & ~amask[r]);
                    if (AFFINE) ix = (ix & amask[r]) | (ux & ~amask[r]);
                }
                dg = left;
                uh = hv;
                ux = ix;
                h[r] = hv;
            }
            diag_top = up_h;
            send_h = uh;
            send_x = ux;
        }
        if (!CHECK) {
            if (lane == 31) {
                smOutH[(s + 1) & 31] = (int)send_h;
                if (AFFINE) smOutX[(s + 1) & 31] = (int)send_x;
            }
            if (s == 30) flush(t0);
        } else {
            const int j31 = t0 + s - 30;
            if (j31 >= 1 && j31 <= n) {
                if (lane == 31) {
                    smOutH[(j31 - 1) & 31] = (int)send_h;
                    if (AFFINE) smOutX[(j31 - 1) & 31] = (int)send_x;
                }
                if (((j31 - 1) & 31) == 31 || j31 == n) flush(j31);
            }
        }
    };
    if (rb > 0) {
        wait_for(min(32, n));
        const int jj = 1 + lane;
        if (jj <= n) {
            nHf = ls_ldcg(inHf + jj);
            nHr = ls_ldcg(inHr + jj);
            if (AFFINE) {
                nXf = ls_ldcg(inXf + jj);
                nXr = ls_ldcg(inXr + jj);
            }
        }
    }
    smT[32 + lane] = make_uint2(0x80808080u, 0x80808080u);
    if (lane == 31) {
        sm31[0] = sm31[2] = basef;
        sm31[1] = sm31[3] = baser;
    }
    for (int t0 = 0; t0 < nsteps; t0 += 32) {
        __syncwarp();
        if (t0 > 0) { // re-base: my first row's value becomes 0
            const unsigned d = h[0];
#pragma unroll
            for (int r = 0; r < R; r++) {
                h[r] = __vsub2(h[r], d);
                if (AFFINE) f[r] = __vsub2(f[r], d);
            }
            diag_top = __vsub2(diag_top, d);
            send_h = __vsub2(send_h, d);
            if (AFFINE) send_x = __vsub2(send_x, d);
            if (lane == 31) {
                sm31[2] = basef;
                sm31[3] = baser;
            }
            basef += ls2_lo(d);
            baser += ls2_hi(d);
            if (lane == 31) {
                sm31[0] = basef;
                sm31[1] = baser;
            }
        }
        {
            const int pf = __shfl_up_sync(SEQA_FULL, basef, 1), pr = __shfl_up_sync(SEQA_FULL, baser, 1);
            adj = ls2_pack(pf - basef, pr - baser); // the lane above me -> my frame
            const int b0f = __shfl_sync(SEQA_FULL, basef, 0), b0r = __shfl_sync(SEQA_FULL, baser, 0);
            const int jj = t0 + 1 + lane; // stage the row above and the profiles of columns t0+1 .. t0+32
            const int hf = rb > 0 ? nHf : border_hrow(bf, jj), hr = rb > 0 ? nHr : border_hrow(br, jj);
            smInH[lane] = (int)ls2_pack(hf + gogo - b0f - cnorm * jj, hr + gogo - b0r - cnorm * jj);
            if (AFFINE) {
                const int xf = rb > 0 ? nXf : bf.ixA + jj * bf.ixB, xr = rb > 0 ? nXr : br.ixA + jj * br.ixB;
                smInX[lane] = (int)ls2_pack(xf - b0f, xr - b0r);
            }
            const uint2 prevT = smT[32 + lane];
            smT[lane] = prevT;
            uint2 T = make_uint2(0x80808080u, 0x80808080u); // beyond the matrix: every score -128
            if (jj <= n) {
                T.x = mm4 ^ (mx << (8 * pk_code(b[jj - 1])));
                T.y = mm4 ^ (mx << (8 * pk_code(b[n - jj])));
            }
            smT[32 + lane] = T;
            const int base = t0 + 32;
            if (rb > 0 && base < n) {
                wait_for(min(base + 32, n));
                if (jj + 32 <= n) {
                    nHf = ls_ldcg(inHf + jj + 32);
                    nHr = ls_ldcg(inHr + jj + 32);
                    if (AFFINE) {
                        nXf = ls_ldcg(inXf + jj + 32);
                        nXr = ls_ldcg(inXr + jj + 32);
                    }
                }
            }
        }
        __syncwarp();
        if (t0 >= 32 && t0 + 32 <= n) {
#pragma unroll 4
            for (int s = 0; s < 32; s++) step(std::false_type(), s, t0);
        } else {
            const int send = min(32, nsteps - t0);
            for (int s = 0; s < send; s++) step(std::true_type(), s, t0);
        }
    }
    __syncwarp();
}

template <bool MM, int R>
__device__ __forceinline__ void ls_run_task2(const LsArgs &A, const LsSweep &S, const LsNode &nd, int rb, int *sm)
{
    const uint32_t p = (uint32_t)nd.pair;
    const int nblk = S.nblk & 0x0fffffff;
    const int mid = nd.m / 2, mf = mid, mr = nd.m - mid; // include/SAHirschberg.h:129, include/SAMyersMiller.h:164
    const uint8_t *a = A.bases + A.off1[p] + nd.i0;
    const uint8_t *b = A.bases + A.off2[p] + nd.j0;
    const int n = nd.n;
    const uint64_t w = A.row_w[p];
    int *base = A.rows + A.row_off[p] + nd.j0 + nd.q;
    Borders bf, br;
    if (!MM) { // NWScore borders, include/SAHirschberg.h:25-29,37
        bf.hcolA = 0; bf.hcolB = A.sc.gap; bf.hrowA = 0; bf.hrowB = A.sc.gap;
        bf.ixA = bf.ixB = bf.iyA = bf.iyB = 0;
        br = bf;
    } else { // include/SAMyersMiller.h:172-198 (forward, tb) / :247-270 (reverse, te)
        const int g = A.sc.go, hh = A.sc.ge;
        bf.hcolA = nd.tb; bf.hcolB = hh; bf.hrowA = g; bf.hrowB = hh;
        bf.ixA = 2 * g; bf.ixB = hh; bf.iyA = nd.tb + g; bf.iyB = hh;
        br = bf;
        br.hcolA = nd.te; br.iyA = nd.te + g;
    }
    const bool last = rb == nblk - 1;
    const int pi = (rb - 1) & 1, po = rb & 1;
    const int *inHf = base + (uint64_t)(LsArr<MM>::BND_F + pi) * w, *inHr = base + (uint64_t)(LsArr<MM>::BND_R + pi) * w;
    const int *inXf = base + (uint64_t)(LsArr<MM>::BND_F + 2 + pi) * w, *inXr = base + (uint64_t)(LsArr<MM>::BND_R + 2 + pi) * w;
    int *outHf = base + (uint64_t)(last ? LsArr<MM>::LAST_H_F : LsArr<MM>::BND_F + po) * w;
    int *outHr = base + (uint64_t)(last ? LsArr<MM>::LAST_H_R : LsArr<MM>::BND_R + po) * w;
    int *outXf = base + (uint64_t)(last ? LsArr<MM>::LAST_X_F : LsArr<MM>::BND_F + 2 + po) * w;
    int *outXr = base + (uint64_t)(last ? LsArr<MM>::LAST_X_R : LsArr<MM>::BND_R + 2 + po) * w;
    const int *in_prog = A.prog + S.task0 + (rb > 0 ? rb - 1 : 0);
    int *out_prog = last ? nullptr : A.prog + S.task0 + rb;
    const bool full = (rb + 1) * 32 * R <= mf; // mf <= mr
    if (full)
        ls_block2<MM, R, false>(A.sc, bf, br, a, mf, a + mid, mr, b, n, rb, inHf, inHr, inXf, inXr, in_prog, outHf, outHr, outXf,
                                outXr, out_prog, sm);
    else
        ls_block2<MM, R, true>(A.sc, bf, br, a, mf, a + mid, mr, b, n, rb, inHf, inHr, inXf, inXr, in_prog, outHf, outHr, outXf,
                               outXr, out_prog, sm);
}

// resident CTAs per SM the two variants are compiled for (register caps 65536 / (LS_BLOCK * n)); measured, DESIGN.md 4.5
#ifndef LS2_MINB_HB
#define LS2_MINB_HB 6
#endif
#ifndef LS2_MINB_MM
#define LS2_MINB_MM 4
#endif
template <bool MM>
__global__ void __launch_bounds__(LS_BLOCK, MM ? LS2_MINB_MM : LS2_MINB_HB) ls_sweep2_kernel(LsArgs A)
{
    __shared__ uint2 smem[LS_BLOCK / 32][LS2_SMEM_INTS / 2]; // uint2: the profile ring needs 8-byte alignment
    const int lane = threadIdx.x & 31;
    int *sm = reinterpret_cast<int *>(smem[threadIdx.x >> 5]);
    const uint32_t ntasks = A.cnt[2];
    if (*A.overflow) return;
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(&A.cnt[3], 1u);
        g = __shfl_sync(SEQA_FULL, g, 0);
        if (g >= ntasks) break;
        const LsTask T = A.tasks[g];
        const LsSweep S = A.sweeps[T.sweep];
        const LsNode nd = A.in[S.node];
        switch ((S.nblk >> 28) & 7) {
        case 0: ls_run_task2<MM, 1>(A, S, nd, (int)T.rb, sm); break;
        case 1: ls_run_task2<MM, 2>(A, S, nd, (int)T.rb, sm); break;
        case 2: ls_run_task2<MM, 4>(A, S, nd, (int)T.rb, sm); break;
        default: ls_run_task2<MM, 8>(A, S, nd, (int)T.rb, sm); break;
        }
    }
}
