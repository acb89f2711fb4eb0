// Generic int32 warp-wavefront DP engine + the fill / walk kernels built on it.
//
// One warp owns one DP sweep.  Rows are cut into blocks of 32*R; inside a block lane l owns the R
// consecutive rows [blk*32R + l*R, +R) and keeps their "left" state in registers.  At step t lane l
// processes column j = t - l + 1 (anti-diagonal wavefront): the H (and Ix) value of the row above its
// strip arrives from lane l-1 by __shfl_up_sync; lane 0 takes it from the boundary row the previous row
// block left in global memory (prefetched 32 columns at a time, coalesced); lane 31 writes the boundary
// row for the next block (buffered, coalesced).  Cell order differs from the reference's row-major order
// but every cell only reads its three finished neighbours, so values are identical; the one
// order-dependent quantity, the local-alignment end cell ("last maximum in row-major order",
// reference include/SASmithWaterman.h:177), is reduced with the key (score, i, j).
//
// Traceback directions are derived per cell with the reference's own equality tests and priority
// (diag > up > left, include/SANeedlemanWunsch.h:171-229; Gotoh: diag > Ix > Iy and extend before
// open, include/SAGlobalGotoh.h:260-419) and packed 2 bits (linear) / 4 bits (affine) per cell.
#pragma once
#include "seqa_common.cuh"

// ---- direction-matrix layout ------------------------------------------------------------------------
// word(i,j) = ((blk*njb + jb)*32 + lane)*R + r with i-1 = blk*32R + lane*R + r, jb = (j-1)/CPW.
// A lane's R words of one column block are contiguous and adjacent lanes are adjacent, so the L2 merges
// the staggered stores of a warp into full lines.
template <bool AFFINE> struct DirFmt {
    static constexpr int BITS = AFFINE ? 4 : 2;
    static constexpr int CPW = 32 / BITS; // cells per 32-bit word
    static constexpr unsigned MASK = (1u << BITS) - 1u;
};

__host__ __device__ inline uint64_t dir_words(int m, int n, int R, bool affine)
{
    if (m <= 0 || n <= 0) return 0;
    const int cpw = affine ? 8 : 16;
    const uint64_t nblk = ((uint64_t)m + 32 * R - 1) / (32 * R);
    const uint64_t njb = ((uint64_t)n + cpw - 1) / cpw;
    return nblk * njb * 32ull * (uint64_t)R;
}

template <bool AFFINE>
__device__ __forceinline__ unsigned dir_get(const uint32_t *dir, int i, int j, int R, int njb)
{
    const int ii = i - 1, jj = j - 1;
    const int rpb = 32 * R;
    const int blk = ii / rpb, rem = ii - blk * rpb;
    const int lane = rem / R, r = rem - lane * R;
    const int jb = jj / DirFmt<AFFINE>::CPW;
    const uint32_t w = dir[((size_t)(blk * njb + jb) * 32 + lane) * R + r];
    return (w >> (DirFmt<AFFINE>::BITS * (jj % DirFmt<AFFINE>::CPW))) & DirFmt<AFFINE>::MASK;
}

struct WfResult {
    int corner;              // H(m,n)
    int best_s, best_i, best_j; // local only
};

// The engine.  Preconditions: m >= 1, n >= 1; all lanes of the warp call it with identical arguments.
template <bool AFFINE, bool LOCAL, bool DIRS, int R>
__device__ __forceinline__ WfResult wavefront(const DevScoring &sc, const Borders &bd,
                                              const uint8_t *__restrict__ a, int m,
                                              const uint8_t *__restrict__ b, int n, bool rev,
                                              uint32_t *__restrict__ dir,
                                              int *__restrict__ boundH, int *__restrict__ boundX,
                                              int *__restrict__ lastH, int *__restrict__ lastX)
{
    constexpr int BITS = DirFmt<AFFINE>::BITS;
    constexpr int CPW = DirFmt<AFFINE>::CPW;
    const int lane = threadIdx.x & 31;
    const int rpb = 32 * R;
    const int nblk = (m + rpb - 1) / rpb;
    const int njb = (n + CPW - 1) / CPW;
    const int gogo = sc.go + sc.ge;

    WfResult res;
    res.corner = 0;
    res.best_s = INT_MIN;
    res.best_i = 0;
    res.best_j = 0;

    for (int blk = 0; blk < nblk; blk++) {
        const int row0 = blk * rpb + lane * R; // rows above my strip
        const int nact = min(max(m - row0, 0), R);
        const bool last_blk = (blk == nblk - 1);
        uint8_t ab[R];
        int h[R], f[R];
        uint32_t acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = row0 + r + 1;
            ab[r] = (r < nact) ? (rev ? a[m - i] : a[i - 1]) : (uint8_t)0;
            h[r] = bd.hcolA + i * bd.hcolB;
            f[r] = AFFINE ? bd.iyA + i * bd.iyB : 0;
            acc[r] = 0;
        }
        int diag_top = border_hcol(bd, row0);
        int send_h = 0, send_x = 0;
        int bbufH = 0, bbufX = 0, wbufH = 0, wbufX = 0;
        const int my_last = m - 1 - row0; // strip index of row m if it is mine (last block only)
        const int nsteps = n + 31;
        for (int t = 0; t < nsteps; t++) {
            if (blk > 0 && (t & 31) == 0) {
                const int jj = t + 1 + lane;
                if (jj <= n) {
                    bbufH = boundH[jj];
                    if (AFFINE) bbufX = boundX[jj];
                }
            }
            int up_h = __shfl_up_sync(SEQA_FULL, send_h, 1);
            int up_x = AFFINE ? __shfl_up_sync(SEQA_FULL, send_x, 1) : 0;
            const int j = t - lane + 1;
            if (blk > 0) {
                const int bh = __shfl_sync(SEQA_FULL, bbufH, t & 31);
                const int bx = AFFINE ? __shfl_sync(SEQA_FULL, bbufX, t & 31) : 0;
                if (lane == 0) {
                    up_h = bh;
                    up_x = bx;
                }
            } else if (lane == 0) {
                up_h = border_hrow(bd, j);
                up_x = AFFINE ? bd.ixA + j * bd.ixB : 0;
            }
            if (j >= 1 && j <= n) {
                const uint8_t bj = rev ? b[n - j] : b[j - 1];
                int dg = diag_top, uh = up_h, ux = up_x;
                const int sh = BITS * ((j - 1) % CPW);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (r < nact) {
                        const int left = h[r];
                        const bool eq = (ab[r] == bj);
                        const int d = diag_cand(sc, dg, eq);
                        int hv;
                        unsigned code;
                        if (!AFFINE) {
                            const int u = uh + sc.gap, l = left + sc.gap;
                            hv = max(max(d, u), l);
                            if (LOCAL) {
                                hv = max(hv, 0);
                                code = (hv == 0) ? 0u : (hv == d ? 1u : (hv == u ? 2u : 3u));
                            } else {
                                code = (hv == d) ? 1u : (hv == u ? 2u : 3u);
                            }
                        } else {
                            const int ixe = ux + sc.ge, ix = max(uh + gogo, ixe);
                            const int iye = f[r] + sc.ge, iy = max(left + gogo, iye);
                            hv = max(max(d, ix), iy);
                            if (LOCAL) {
                                hv = max(hv, 0);
                                const int s0 = max(d, 0);
                                code = (hv == s0) ? (s0 <= 0 ? 0u : 1u) : (hv == ix ? 2u : 3u);
                            } else {
                                code = (hv == d) ? 1u : (hv == ix ? 2u : 3u);
                            }
                            code |= (ix == ixe ? 4u : 0u) | (iy == iye ? 8u : 0u);
                            f[r] = iy;
                            ux = ix;
                        }
                        if (LOCAL) {
                            const int i = row0 + r + 1;
                            if (hv > res.best_s ||
                                (hv == res.best_s && (i > res.best_i || (i == res.best_i && j > res.best_j)))) {
                                res.best_s = hv;
                                res.best_i = i;
                                res.best_j = j;
                            }
                        }
                        if (DIRS) acc[r] |= code << sh;
                        dg = left;
                        uh = hv;
                        h[r] = hv;
                    }
                }
                diag_top = up_h;
                send_h = h[R - 1];
                send_x = ux;
                if (DIRS && (((j - 1) % CPW) == CPW - 1 || j == n)) {
                    const size_t base = ((size_t)(blk * njb + (j - 1) / CPW) * 32 + lane) * R;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        if (r < nact) dir[base + r] = acc[r];
                        acc[r] = 0;
                    }
                }
                if (last_blk && my_last >= 0 && my_last < R) {
                    int hm = h[0], xm = ux;
#pragma unroll
                    for (int r = 1; r < R; r++)
                        if (r == my_last) hm = h[r];
                    if (lastH) {
                        lastH[j] = hm;
                        if (AFFINE) lastX[j] = xm;
                    }
                    if (j == n) res.corner = hm;
                }
            }
            if (!last_blk) {
                const int j31 = t - 30;
                const int vh = __shfl_sync(SEQA_FULL, send_h, 31);
                const int vx = AFFINE ? __shfl_sync(SEQA_FULL, send_x, 31) : 0;
                if (j31 >= 1 && j31 <= n) {
                    const int slot = (j31 - 1) & 31;
                    if (lane == slot) {
                        wbufH = vh;
                        wbufX = vx;
                    }
                    if (slot == 31 || j31 == n) {
                        const int base = j31 - slot;
                        if (lane <= slot) {
                            boundH[base + lane] = wbufH;
                            if (AFFINE) boundX[base + lane] = wbufX;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    // broadcast H(m,n) from the lane that owns row m
    const int owner = ((m - 1) % rpb) / R;
    res.corner = __shfl_sync(SEQA_FULL, res.corner, owner);
    if (LOCAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int s = __shfl_xor_sync(SEQA_FULL, res.best_s, o);
            const int i = __shfl_xor_sync(SEQA_FULL, res.best_i, o);
            const int j = __shfl_xor_sync(SEQA_FULL, res.best_j, o);
            if (s > res.best_s || (s == res.best_s && (i > res.best_i || (i == res.best_i && j > res.best_j)))) {
                res.best_s = s;
                res.best_i = i;
                res.best_j = j;
            }
        }
    }
    return res;
}

// ---- batch fill kernel -------------------------------------------------------------------------------
struct FillArgs {
    const uint8_t *bases;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const uint32_t *idx;     // pair ids of this launch (position k -> pair idx[k])
    uint64_t count;
    const uint64_t *dir_off; // word offset of each pair's direction matrix inside dir, indexed by position k
    uint32_t *dir;
    int *bound;              // per-warp boundary scratch, 2*(bound_stride) ints per warp
    int bound_stride;
    int32_t *score;
    uint32_t *end_i, *end_j;
    DevScoring sc;
    Borders bd;
};

template <bool AFFINE, bool LOCAL, int R>
__global__ void __launch_bounds__(128) fill_i32_kernel(FillArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int *boundH = A.bound + gw * 2 * (uint64_t)A.bound_stride;
    int *boundX = boundH + A.bound_stride;
    for (uint64_t k = gw; k < A.count; k += nw) {
        const uint64_t p = A.idx[k];
        const int m = (int)A.len1[p], n = (int)A.len2[p];
        if (m == 0 || n == 0) {
            // reference: NW/Gotoh borders only (include/SANeedlemanWunsch.h:59-62); SW with an empty input
            // starts and ends at (0,0) (include/SASmithWaterman.h:234-238, fresh object)
            if (lane == 0) {
                A.score[p] = LOCAL ? 0 : (m == 0 ? border_hrow(A.bd, n) : border_hcol(A.bd, m));
                A.end_i[p] = LOCAL ? 0u : (uint32_t)m;
                A.end_j[p] = LOCAL ? 0u : (uint32_t)n;
            }
            continue;
        }
        WfResult r = wavefront<AFFINE, LOCAL, true, R>(A.sc, A.bd, A.bases + A.off1[p], m, A.bases + A.off2[p], n, false,
                                                      A.dir + A.dir_off[k], boundH, boundX, nullptr, nullptr);
        if (lane == 0) {
            A.score[p] = LOCAL ? r.best_s : r.corner;
            A.end_i[p] = LOCAL ? (uint32_t)r.best_i : (uint32_t)m;
            A.end_j[p] = LOCAL ? (uint32_t)r.best_j : (uint32_t)n;
        }
    }
}

// ---- traceback walk ----------------------------------------------------------------------------------
struct WalkArgs {
    const uint32_t *len1, *len2;
    const uint32_t *idx;
    uint64_t count;
    const uint64_t *dir_off;
    const uint32_t *dir;
    int R;
    const uint32_t *end_i, *end_j;
    uint32_t *start_i, *start_j;
    uint8_t *slots;           // per-pair op slots of len1+len2 bytes
    const uint64_t *slot_off; // byte offset of each pair's slot
    uint32_t *slot_start;     // first used byte inside the slot (ops are written back-to-front)
    uint32_t *ops_len;
};

// One thread per pair follows the stored codes from the end cell.  State machine = reference buildResult:
// linear include/SANeedlemanWunsch.h:155-231, include/SASmithWaterman.h:220-339; affine
// include/SAGlobalGotoh.h:235-422, include/SALocalGotoh.h:275-473.
template <bool AFFINE, bool LOCAL>
__global__ void walk_kernel(WalkArgs A)
{
    const uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= A.count) return;
    const uint64_t p = A.idx[k0];
    const int m = (int)A.len1[p], n = (int)A.len2[p];
    const int R = A.R;
    const int njb = (n + DirFmt<AFFINE>::CPW - 1) / DirFmt<AFFINE>::CPW;
    const uint32_t *dir = A.dir + A.dir_off[k0];
    uint8_t *slot = A.slots + A.slot_off[p];
    int k = m + n;
    int i = (int)A.end_i[p], j = (int)A.end_j[p];
    int state = 0;
    if (!LOCAL) {
        while (i > 0 || j > 0) {
            if (AFFINE ? (j == 0) : (i != 0 && j == 0)) { // column 0: up
                slot[--k] = 1; i--; continue;
            }
            if (i == 0) { // row 0: left
                slot[--k] = 2; j--; continue;
            }
            const unsigned c = dir_get<AFFINE>(dir, i, j, R, njb);
            if (!AFFINE) {
                if (c == 1u) { slot[--k] = 0; i--; j--; }
                else if (c == 2u) { slot[--k] = 1; i--; }
                else { slot[--k] = 2; j--; }
            } else {
                if (state == 0) {
                    const unsigned hs = c & 3u;
                    if (hs == 1u) { slot[--k] = 0; i--; j--; }
                    else state = (hs == 2u) ? 1 : 2; // switch matrix, no move (include/SAGlobalGotoh.h:355-363,411-419)
                } else if (state == 1) {
                    slot[--k] = 1; i--;
                    if (!(c & 4u)) state = 0; // opened here (:344-354); else extension (:336-343)
                } else {
                    slot[--k] = 2; j--;
                    if (!(c & 8u)) state = 0;
                }
            }
        }
    } else {
        while (i > 0 && j > 0) {
            const unsigned c = dir_get<AFFINE>(dir, i, j, R, njb);
            if (!AFFINE) {
                if (c == 0u) break;
                if (c == 1u) { slot[--k] = 0; i--; j--; }
                else if (c == 2u) { slot[--k] = 1; i--; }
                else { slot[--k] = 2; j--; }
            } else {
                if (state == 0) {
                    const unsigned hs = c & 3u;
                    if (hs == 0u) break;
                    if (hs == 1u) { slot[--k] = 0; i--; j--; }
                    else state = (hs == 2u) ? 1 : 2;
                } else if (state == 1) {
                    slot[--k] = 1; i--;
                    if (!(c & 4u)) state = 0;
                } else {
                    slot[--k] = 2; j--;
                    if (!(c & 8u)) state = 0;
                }
            }
        }
    }
    A.start_i[p] = (uint32_t)i;
    A.start_j[p] = (uint32_t)j;
    A.slot_start[p] = (uint32_t)k;
    A.ops_len[p] = (uint32_t)(m + n - k);
}

