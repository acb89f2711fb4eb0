// Linear-space global aligners on the GPU: HirschbergSA (reference include/SAHirschberg.h:11-184) and
// MyersMillerSA (reference include/SAMyersMiller.h:43-420).
//
// The reference recursion is run level by level.  Per recursion depth three kernels run over every live
// sub-problem ("node") of every pair of the batch:
//
//   ls_expand_kernel  one warp per node: leaves write their ops; an internal node becomes two score-only SWEEPS
//                     (forward over its top rows, reverse over its bottom rows), each cut into ROW BLOCKS of 32*R
//                     rows; every (sweep, row block) is one TASK appended to the level's task list.
//   ls_sweep_kernel   persistent: warps draw tasks from an atomic ticket counter.  A task runs the anti-diagonal
//                     warp wavefront (boundary H / gap-state values by __shfl_up_sync) over ALL columns of its
//                     row block; its bottom row streams to the task of the next row block through a global
//                     (L2-resident) boundary row guarded by a progress counter, so the row blocks of one sweep
//                     form a software pipeline across warps and SMs: a 100 kbp x 50 kbp sweep is ~200 warps
//                     deep.  Tickets are handed out in an order in which a task's producer always holds an
//                     earlier ticket, so the producer is running or done whenever its consumer waits: no deadlock
//                     without any residency assumption.  The last row block writes the sweep's final row(s).
//   ls_split_kernel   one warp per internal node: the reference's split rule (with its tie rule) over the final
//                     rows of the two sweeps; pushes the two children onto the next level's node list.
//
// A leaf writes its ops straight into the pair's op slot at position i0 + j0 (at most i0 + j0 ops precede a
// node that starts at cell (i0, j0) and a node covering m rows and n columns emits at most m + n ops, so leaves
// never collide); a final pass squeezes out the unused slot bytes and scores the alignment.
//
// Bit-exactness notes: the recursion is followed all the way down to the
// reference's own leaves -- never short-circuited into a full-matrix aligner, because HirschbergSA's split
// search skips column N (include/SAHirschberg.h:141) and MyersMillerSA's M == 1 leaf is not the textbook one
// (include/SAMyersMiller.h:75-160): both are sub-optimal in a specific way that has to be reproduced.
#pragma once
#include <type_traits>
#include <vector>
#include "seqa_common.cuh"
#include "seqa_wavefront.cuh"

#define LS_HOLE 0xffu
#define LS_MAXR 8                 /* rows per lane of a full row block */
#define LS_NEG (-(1 << 30))       /* "never" diagonal candidate (AllowMismatch == false); |H| < 2^29 is checked on the host */
#define LS_BLOCK 128              /* threads per CTA of the sweep kernel */

// One score-only sweep of an internal node.
struct LsSweep {
    int node;       // index into the level's node list
    int rows;       // rows of the sweep; bit 31 set = reverse sweep (bottom rows, both sequences reversed)
    int nblk;       // row blocks; bits 28-30 = log2(R)
    uint32_t task0; // progress counters of its row blocks live at prog[task0 + rb]
};
struct LsTask {
    uint32_t sweep, rb;
};

struct LsArgs {
    const uint8_t *bases;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const LsNode *in;
    uint32_t n_in;
    LsNode *out;
    uint32_t out_cap;
    uint32_t *cnt;           // [0] nodes pushed to `out`, [1] sweeps, [2] tasks, [3] ticket
    int *overflow;
    LsSweep *sweeps;
    LsTask *tasks;
    uint32_t task_cap;
    int *prog;               // per task: columns of its bottom row published so far
    int *rows;               // scratch rows
    const uint64_t *row_off; // per pair: int offset of its arrays
    const uint32_t *row_w;   // per pair: stride of one array
    uint8_t *slots;
    const uint64_t *slot_off;
    DevScoring sc;
    int force_r;             // > 0: use this many rows per lane everywhere (testing: many row blocks on short pairs)
    int packed;              // forward + reverse sweep of a node as one s16x2 wavefront (seqa_linspace_packed.cuh)
};

#ifdef SEQA_EMU
static inline int ls_ld_volatile(const int *p) { return *(const volatile int *)p; }
static inline void ls_st_volatile(int *p, int v) { *(volatile int *)p = v; }
static inline int ls_ldcg(const int *p) { return *p; }
static inline void ls_pause() { emu_yield(); }
#else
__device__ __forceinline__ int ls_ld_volatile(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void ls_st_volatile(int *p, int v) { *(volatile int *)p = v; }
__device__ __forceinline__ int ls_ldcg(const int *p) { return __ldcg(p); } // L2 only: written by another SM
__device__ __forceinline__ void ls_pause() { __nanosleep(64); }
#endif

__device__ __forceinline__ void ls_emit_run(uint8_t *slot, int from, int count, uint8_t op, int lane)
{
    for (int k = lane; k < count; k += 32) slot[from + k] = op;
}

// ---- Hirschberg ---------------------------------------------------------------------------------------------
// Leaf with a single row or a single column = NeedlemanWunschSA on the views (include/SAHirschberg.h:119-126),
// evaluated in closed form: the traceback of a 1 x n (m x 1) matrix walks left (up) from the corner until the
// first cell where the reference's test order (diag, then up, else left; include/SANeedlemanWunsch.h:171-229)
// leaves the row (column), after which only border moves remain.
__device__ void hb_leaf_thin(const DevScoring &sc, const uint8_t *a, int m, const uint8_t *b, int n, uint8_t *slot, int lane)
{
    const int g = sc.gap;
    if (m == 1) {
        // H[1][j] = max(D_j, U_j, H[1][j-1]+g), D_j = (j-1)g + sim_j, U_j = (j+1)g
        int stop = 0, kind = 0; // kind 1 = diag, 2 = up
        if (lane == 0) {
            int h = g; // H[1][0]
            for (int j = 1; j <= n; j++) {
                const int d = diag_cand(sc, (j - 1) * g, a[0] == b[j - 1]);
                const int u = (j + 1) * g;
                h = max(max(d, u), h + g);
                if (h == d) { stop = j; kind = 1; }
                else if (h == u) { stop = j; kind = 2; }
            }
        }
        stop = __shfl_sync(SEQA_FULL, stop, 0);
        kind = __shfl_sync(SEQA_FULL, kind, 0);
        if (kind == 1) { // LEFT x (stop-1), DIAG, LEFT x (n-stop)
            ls_emit_run(slot, 0, stop - 1, 2, lane);
            if (lane == 0) slot[stop - 1] = 0;
            ls_emit_run(slot, stop, n - stop, 2, lane);
        } else if (kind == 2) { // LEFT x stop, UP, LEFT x (n-stop)
            ls_emit_run(slot, 0, stop, 2, lane);
            if (lane == 0) slot[stop] = 1;
            ls_emit_run(slot, stop + 1, n - stop, 2, lane);
        } else { // reached column 0 in row 1: UP, then LEFT x n
            if (lane == 0) slot[0] = 1;
            ls_emit_run(slot, 1, n, 2, lane);
        }
    } else { // n == 1
        int stop = 0, kind = 0; // kind 1 = diag, 3 = left
        if (lane == 0) {
            int h = g; // H[0][1]
            for (int i = 1; i <= m; i++) {
                const int d = diag_cand(sc, (i - 1) * g, a[i - 1] == b[0]);
                const int u = h + g;
                const int l = (i + 1) * g; // H[i][0] + g
                h = max(max(d, u), l);
                if (h == d) { stop = i; kind = 1; }
                else if (h != u) { stop = i; kind = 3; }
            }
        }
        stop = __shfl_sync(SEQA_FULL, stop, 0);
        kind = __shfl_sync(SEQA_FULL, kind, 0);
        if (kind == 1) { // UP x (stop-1), DIAG, UP x (m-stop)
            ls_emit_run(slot, 0, stop - 1, 1, lane);
            if (lane == 0) slot[stop - 1] = 0;
            ls_emit_run(slot, stop, m - stop, 1, lane);
        } else if (kind == 3) { // UP x stop, LEFT, UP x (m-stop)
            ls_emit_run(slot, 0, stop, 1, lane);
            if (lane == 0) slot[stop] = 2;
            ls_emit_run(slot, stop + 1, m - stop, 1, lane);
        } else { // reached row 0 in column 1: LEFT, then UP x m
            if (lane == 0) slot[0] = 2;
            ls_emit_run(slot, 1, m, 1, lane);
        }
    }
}


// ---- one row block of a score-only sweep -----------------------------------------------------------------------
// Rows [rb*32R, rb*32R + 32R) of an m x n sweep, all n columns, on one warp.  Lane l owns R consecutive rows; at
// step t it computes column t - l + 1.  Columns are taken in chunks of 32 steps.  At the start of a chunk the warp
// stages, in shared memory, the row above the block for the chunk's 32 columns (from the border for rb == 0, else
// from `inH/inX`, published by the task of row block rb-1 and fetched one chunk ahead) and the chunk's 32 column
// symbols (one coalesced load); lane 0 reads its upper neighbour from there, every lane its column symbol.  Lane
// 31's bottom row leaves through shared memory in coalesced 32-column pieces, followed by the progress counter.
// Chunks in which every lane is inside the matrix for all 32 steps run a branch-free step.
// PARTIAL: rows >= m pass the value above them through unchanged, so the sweep's last row arrives at lane 31 like
// any other bottom row.
#define LS_SMEM_INTS 192 /* per warp: inH[32] inX[32] outH[32] outX[32] sym[64] */

template <bool AFFINE, int R, bool PARTIAL>
__device__ __forceinline__ void ls_block(const DevScoring &sc, const Borders &bd, const uint8_t *__restrict__ a, int m,
                                         const uint8_t *__restrict__ b, int n, bool rev, int rb,
                                         const int *inH, const int *inX, const int *in_prog,
                                         int *outH, int *outX, int *out_prog, int *sm)
{
    const int lane = threadIdx.x & 31;
    const int row0 = rb * (32 * R) + lane * R; // rows above my strip
    const int nact = PARTIAL ? min(max(m - row0, 0), R) : R;
    const int gap = sc.gap, ge = sc.ge, gogo = sc.go + sc.ge, match = sc.match;
    const int simx = sc.allow ? sc.mismatch : LS_NEG;
    int ab[R], h[R], f[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = row0 + r + 1;
        ab[r] = (r < nact) ? (int)(rev ? a[m - i] : a[i - 1]) : 0x100; // 0x100 never equals a symbol
        h[r] = bd.hcolA + i * bd.hcolB;
        f[r] = AFFINE ? bd.iyA + i * bd.iyB : 0;
    }
    int diag_top = border_hcol(bd, row0);
    int send_h = 0, send_x = 0;
    int *smInH = sm, *smInX = sm + 32, *smOutH = sm + 64, *smOutX = sm + 96, *smSym = sm + 128;
    const int *mySym = smSym + 32 - lane; // mySym[s] = symbol of column t0 + s - lane + 1
    const int nsteps = n + 31;
    int nextH = 0, nextX = 0, known = 0;
    auto wait_for = [&](int need) { // until the producer has published `need` columns
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
    auto flush = [&](int j31) { // columns ((j31-1) & ~31) + 1 .. j31 of the bottom row leave the warp
        __syncwarp();
        const int base = (j31 - 1) & ~31;
        if (lane < j31 - base) {
            outH[base + 1 + lane] = smOutH[lane];
            if (AFFINE) outX[base + 1 + lane] = smOutX[lane];
        }
        if (out_prog) {
            __threadfence();
            __syncwarp();
            if (lane == 0) ls_st_volatile(out_prog, j31);
        } else {
            __syncwarp();
        }
    };
    // one anti-diagonal step; CHECK = lanes may be outside the matrix (first / last chunks)
    auto step = [&](auto chk, int s, int t0) {
        constexpr bool CHECK = decltype(chk)::value;
        int up_h = __shfl_up_sync(SEQA_FULL, send_h, 1);
        int up_x = AFFINE ? __shfl_up_sync(SEQA_FULL, send_x, 1) : 0;
        const int bh = smInH[s], bx = AFFINE ? smInX[s] : 0;
        if (lane == 0) {
            up_h = bh;
            up_x = bx;
        }
        const int j = t0 + s - lane + 1;
        if (!CHECK || (j >= 1 && j <= n)) {
            const int bj = mySym[s];
            int dg = diag_top, uh = up_h, ux = up_x;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int left = h[r];
                const int sim = (ab[r] == bj) ? match : simx;
                int hv, ix = ux;
                if (!AFFINE) {
                    const int tl = __viaddmax_s32(dg, sim, left + gap); // max(D, L)
                    hv = __viaddmax_s32(uh, gap, tl);                  // max(U, .)
                } else {
                    ix = __viaddmax_s32(uh, gogo, ux + ge);
                    const int iy = __viaddmax_s32(left, gogo, f[r] + ge);
                    hv = __viaddmax_s32(dg, sim, max(ix, iy));
                    f[r] = iy;
                }
                if (PARTIAL && r >= nact) { // below the sweep: hand the last row down
                    hv = uh;
                    ix = ux;
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
        // lane 31 has just finished column j31 = t0 + s - 30; its slot in the outgoing piece is (j31 - 1) & 31
        if (!CHECK) {
            if (lane == 31) {
                smOutH[(s + 1) & 31] = send_h;
                if (AFFINE) smOutX[(s + 1) & 31] = send_x;
            }
            if (s == 30) flush(t0);
        } else {
            const int j31 = t0 + s - 30;
            if (j31 >= 1 && j31 <= n) {
                if (lane == 31) {
                    smOutH[(j31 - 1) & 31] = send_h;
                    if (AFFINE) smOutX[(j31 - 1) & 31] = send_x;
                }
                if (((j31 - 1) & 31) == 31 || j31 == n) flush(j31);
            }
        }
    };
    if (rb > 0) {
        wait_for(min(32, n));
        const int jj = 1 + lane;
        if (jj <= n) {
            nextH = ls_ldcg(inH + jj);
            if (AFFINE) nextX = ls_ldcg(inX + jj);
        }
    }
    smSym[32 + lane] = 0x200;
    for (int t0 = 0; t0 < nsteps; t0 += 32) {
        __syncwarp();
        {
            const int jj = t0 + 1 + lane; // stage the row above and the symbols of columns t0+1 .. t0+32
            smInH[lane] = rb > 0 ? nextH : border_hrow(bd, jj);
            if (AFFINE) smInX[lane] = rb > 0 ? nextX : bd.ixA + jj * bd.ixB;
            const int prev = smSym[32 + lane];
            smSym[lane] = prev;
            smSym[32 + lane] = (jj <= n) ? (int)(rev ? b[n - jj] : b[jj - 1]) : 0x200;
            const int base = t0 + 32;
            if (rb > 0 && base < n) { // boundary of the next chunk, one chunk ahead of its use
                wait_for(min(base + 32, n));
                if (jj + 32 <= n) {
                    nextH = ls_ldcg(inH + jj + 32);
                    if (AFFINE) nextX = ls_ldcg(inX + jj + 32);
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

// scratch arrays of one node (stride w ints): Hirschberg  F, Rv, fwd boundary x2, rev boundary x2           (6)
//                                             MyersMiller CC, DD, RR, SS, fwd (H x2, X x2), rev (H x2, X x2) (12)
template <bool MM> struct LsArr {
    static constexpr int COUNT = MM ? 12 : 6;
    static constexpr int LAST_H_F = 0, LAST_X_F = 1, LAST_H_R = MM ? 2 : 1, LAST_X_R = 3;
    static constexpr int BND_F = MM ? 4 : 2, BND_R = MM ? 8 : 4; // H ping, H pong [, X ping, X pong]
};

template <bool MM, int R>
__device__ __forceinline__ void ls_run_task(const LsArgs &A, const LsSweep &S, const LsNode &nd, int rb, int *sm)
{
    const uint32_t p = (uint32_t)nd.pair;
    const bool rev = S.rows < 0;
    const int rows = S.rows & 0x7fffffff, nblk = S.nblk & 0x0fffffff;
    const int mid = nd.m / 2; // include/SAHirschberg.h:129, include/SAMyersMiller.h:164
    const uint8_t *a = A.bases + A.off1[p] + nd.i0 + (rev ? mid : 0);
    const uint8_t *b = A.bases + A.off2[p] + nd.j0;
    const int n = nd.n;
    const uint64_t w = A.row_w[p];
    int *base = A.rows + A.row_off[p] + nd.j0 + nd.q;
    Borders bd;
    if (!MM) { // NWScore borders, include/SAHirschberg.h:25-29,37
        bd.hcolA = 0; bd.hcolB = A.sc.gap; bd.hrowA = 0; bd.hrowB = A.sc.gap;
        bd.ixA = bd.ixB = bd.iyA = bd.iyB = 0;
    } else { // include/SAMyersMiller.h:172-198 (forward, tb) / :247-270 (reverse, te):
             // H(i,0)=t+i*h, H(0,j)=g+j*h, DD(0,j)=H(0,j)+g, e(i,0)=H(i,0)+g
        const int g = A.sc.go, h = A.sc.ge, t = rev ? nd.te : nd.tb;
        bd.hcolA = t; bd.hcolB = h; bd.hrowA = g; bd.hrowB = h;
        bd.ixA = 2 * g; bd.ixB = h; bd.iyA = t + g; bd.iyB = h;
    }
    const int bnd = rev ? LsArr<MM>::BND_R : LsArr<MM>::BND_F;
    const bool last = rb == nblk - 1;
    const int *inH = base + (uint64_t)(bnd + ((rb - 1) & 1)) * w;
    const int *inX = base + (uint64_t)(bnd + 2 + ((rb - 1) & 1)) * w;
    int *outH = last ? base + (uint64_t)(rev ? LsArr<MM>::LAST_H_R : LsArr<MM>::LAST_H_F) * w
                     : base + (uint64_t)(bnd + (rb & 1)) * w;
    int *outX = last ? base + (uint64_t)(rev ? LsArr<MM>::LAST_X_R : LsArr<MM>::LAST_X_F) * w
                     : base + (uint64_t)(bnd + 2 + (rb & 1)) * w;
    const int *in_prog = A.prog + S.task0 + (rb > 0 ? rb - 1 : 0);
    int *out_prog = last ? nullptr : A.prog + S.task0 + rb;
    const bool full = (rb + 1) * 32 * R <= rows;
    if (full)
        ls_block<MM, R, false>(A.sc, bd, a, rows, b, n, rev, rb, inH, inX, in_prog, outH, outX, out_prog, sm);
    else
        ls_block<MM, R, true>(A.sc, bd, a, rows, b, n, rev, rb, inH, inX, in_prog, outH, outX, out_prog, sm);
}

template <bool MM>
__global__ void __launch_bounds__(LS_BLOCK) ls_sweep_kernel(LsArgs A)
{
    __shared__ int smem[LS_BLOCK / 32][LS_SMEM_INTS];
    const int lane = threadIdx.x & 31;
    int *sm = smem[threadIdx.x >> 5];
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
        case 0: ls_run_task<MM, 1>(A, S, nd, (int)T.rb, sm); break;
        case 1: ls_run_task<MM, 2>(A, S, nd, (int)T.rb, sm); break;
        case 2: ls_run_task<MM, 4>(A, S, nd, (int)T.rb, sm); break;
        default: ls_run_task<MM, 8>(A, S, nd, (int)T.rb, sm); break;
        }
    }
}

// ---- expand: leaves emit their ops, internal nodes become sweeps + tasks ------------------------------------------
__device__ __forceinline__ void ls_push(const LsArgs &A, const LsNode &nd)
{
    const uint32_t k = atomicAdd(&A.cnt[0], 1u);
    if (k < A.out_cap)
        A.out[k] = nd;
    else
        *A.overflow = 1;
}

// rows per lane for a sweep of `rows` rows: the smallest of 1/2/4/8 whose single block covers it, else 8
__device__ __forceinline__ int ls_pick_rsel(const LsArgs &A, int rows)
{
    if (A.force_r > 0) return A.force_r == 1 ? 0 : A.force_r == 2 ? 1 : A.force_r == 4 ? 2 : 3;
    return rows <= 32 ? 0 : rows <= 64 ? 1 : rows <= 128 ? 2 : 3;
}

__device__ __forceinline__ void ls_make_sweeps(const LsArgs &A, uint32_t node, int rows_f, int rows_r, int lane)
{
    if (A.packed) { // one dual sweep: row block rb of both sweeps is one task (rows_f <= rows_r <= rows_f + 1)
        const int rs = ls_pick_rsel(A, rows_r);
        const int nb = (rows_r + (32 << rs) - 1) / (32 << rs);
        uint32_t s0 = 0, t0 = 0;
        if (lane == 0) {
            s0 = atomicAdd(&A.cnt[1], 1u);
            t0 = atomicAdd(&A.cnt[2], (uint32_t)nb);
            if (t0 + (uint32_t)nb > A.task_cap) {
                *A.overflow = 1;
                t0 = 0xffffffffu;
            }
        }
        s0 = __shfl_sync(SEQA_FULL, s0, 0);
        t0 = __shfl_sync(SEQA_FULL, t0, 0);
        if (t0 == 0xffffffffu) return;
        if (lane == 0) {
            LsSweep D;
            D.node = (int)node; D.rows = rows_r; D.nblk = nb | (rs << 28); D.task0 = t0;
            A.sweeps[s0] = D;
        }
        for (int rb = lane; rb < nb; rb += 32) {
            LsTask T; T.sweep = s0; T.rb = (uint32_t)rb;
            A.tasks[t0 + (uint32_t)rb] = T;
        }
        return;
    }
    const int rf = ls_pick_rsel(A, rows_f), rr = ls_pick_rsel(A, rows_r);
    const int nf = (rows_f + (32 << rf) - 1) / (32 << rf), nr = (rows_r + (32 << rr) - 1) / (32 << rr);
    uint32_t s0 = 0, t0 = 0;
    if (lane == 0) {
        s0 = atomicAdd(&A.cnt[1], 2u);
        t0 = atomicAdd(&A.cnt[2], (uint32_t)(nf + nr));
        if (t0 + (uint32_t)(nf + nr) > A.task_cap) {
            *A.overflow = 1; // the sweep / split kernels do nothing once this is set; the host fails the call
            t0 = 0xffffffffu;
        }
    }
    s0 = __shfl_sync(SEQA_FULL, s0, 0);
    t0 = __shfl_sync(SEQA_FULL, t0, 0);
    if (t0 == 0xffffffffu) return;
    if (lane == 0) {
        LsSweep F, Rv;
        F.node = (int)node; F.rows = rows_f; F.nblk = nf | (rf << 28); F.task0 = t0;
        Rv.node = (int)node; Rv.rows = (int)((unsigned)rows_r | 0x80000000u); Rv.nblk = nr | (rr << 28); Rv.task0 = t0 + (uint32_t)nf;
        A.sweeps[s0] = F;
        A.sweeps[s0 + 1] = Rv;
    }
    // task order: forward and reverse row blocks alternate, each chain in ascending row-block order
    for (int rb = lane; rb < nf; rb += 32) {
        LsTask T; T.sweep = s0; T.rb = (uint32_t)rb;
        A.tasks[t0 + (uint32_t)(rb + min(rb, nr))] = T;
    }
    for (int rb = lane; rb < nr; rb += 32) {
        LsTask T; T.sweep = s0 + 1; T.rb = (uint32_t)rb;
        A.tasks[t0 + (uint32_t)(rb + min(rb + 1, nf))] = T;
    }
}

template <bool MM>
__global__ void __launch_bounds__(128) ls_expand_kernel(LsArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = gw; k < A.n_in; k += nw) {
        const LsNode nd = A.in[k];
        const uint32_t p = (uint32_t)nd.pair;
        const uint8_t *a = A.bases + A.off1[p] + nd.i0;
        const uint8_t *b = A.bases + A.off2[p] + nd.j0;
        uint8_t *slot = A.slots + A.slot_off[p] + nd.i0 + nd.j0;
        const int m = nd.m, n = nd.n;
        if (!MM) {
            if (m == 0) { // include/SAHirschberg.h:105
                ls_emit_run(slot, 0, n, 2, lane);
            } else if (n == 0) { // :112
                ls_emit_run(slot, 0, m, 1, lane);
            } else if (m == 1 || n == 1) { // :119
                hb_leaf_thin(A.sc, a, m, b, n, slot, lane);
            } else {
                ls_make_sweeps(A, k, m / 2, m - m / 2, lane); // :129-139
            }
        } else {
            const int g = A.sc.go, h = A.sc.ge, tb = nd.tb, te = nd.te;
            if (n == 0) { // include/SAMyersMiller.h:57
                ls_emit_run(slot, 0, m, 1, lane);
            } else if (m == 0) { // :67
                ls_emit_run(slot, 0, n, 2, lane);
            } else if (m == 1) { // :75-160
                const int mx = max(tb, te) + h + (g + h * n);
                int best = INT_MIN, bj = 0;
                for (int j = 1 + lane; j <= n; j += 32) {
                    const bool eq = a[0] == b[j - 1];
                    int v = mx;
                    if (A.sc.allow || eq) v = max(g + h * (j - 1) + (eq ? A.sc.match : A.sc.mismatch) + g + h * (n - j), mx); // :104-113
                    if (v > best) { best = v; bj = j; } // first strictly greatest (:121); lanes ascend in j
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int ov = __shfl_xor_sync(SEQA_FULL, best, o);
                    const int oj = __shfl_xor_sync(SEQA_FULL, bj, o);
                    if (oj != 0 && (bj == 0 || ov > best || (ov == best && oj < bj))) { best = ov; bj = oj; }
                }
                const bool split = !A.sc.allow && a[0] != b[bj - 1]; // :142-147: (a,-) then (-,b)
                ls_emit_run(slot, 0, bj - 1, 2, lane);
                if (lane == 0) {
                    if (split) { slot[bj - 1] = 1; slot[bj] = 2; }
                    else slot[bj - 1] = 0;
                }
                ls_emit_run(slot, bj + (split ? 1 : 0), n - bj, 2, lane);
            } else {
                ls_make_sweeps(A, k, m / 2, m - m / 2, lane); // :164, :190, :272
            }
        }
    }
}

// ---- split: the reference's midpoint rule over the final rows of the two sweeps ------------------------------------
template <bool MM>
__global__ void __launch_bounds__(128) ls_split_kernel(LsArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t per = A.packed ? 1u : 2u; // sweep entries per internal node
    const uint32_t nint = *A.overflow ? 0u : A.cnt[1] / per; // internal nodes of this level
    for (uint32_t k = gw; k < nint; k += nw) {
        const LsNode nd = A.in[A.sweeps[per * k].node];
        const uint32_t p = (uint32_t)nd.pair;
        const int m = nd.m, n = nd.n, mid = m / 2;
        const uint64_t w = A.row_w[p];
        const int *base = A.rows + A.row_off[p] + nd.j0 + nd.q;
        if (!MM) {
            const int *F = base, *Rv = base + w;
            // Seq2Mid = argmax over i in [0, n-1] of F[i] + Rv[n-i], ties -> largest i (include/SAHirschberg.h:141-149);
            // column 0 of both rows is the border value
            int best = INT_MIN, bi = 0;
            for (int i = lane; i < n; i += 32) {
                const int f = i == 0 ? mid * A.sc.gap : F[i];
                const int v = f + Rv[n - i];
                if (v >= best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(SEQA_FULL, best, o);
                const int oi = __shfl_xor_sync(SEQA_FULL, bi, o);
                if (ov > best || (ov == best && oi > bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                LsNode l = nd, r = nd;
                l.m = mid; l.n = bi; l.q = 2 * nd.q;
                r.i0 = nd.i0 + mid; r.m = m - mid; r.j0 = nd.j0 + bi; r.n = n - bi; r.q = 2 * nd.q + 1;
                ls_push(A, l); // :151-155
                ls_push(A, r); // :157-161
            }
        } else {
            const int g = A.sc.go, h = A.sc.ge;
            const int *CC = base, *DD = base + w, *RR = base + 2 * w, *SS = base + 3 * w;
            const int c0 = nd.tb + mid * h, r0 = nd.te + (m - mid) * h; // CC[0] = DD[0] (:238), RR[0] = SS[0] (:313)
            // midpoint (:315-340): first strictly greatest of max(CC+RR, DD+SS-g) over j = 0..N; RR/SS are stored
            // by reversed column index
            int best = INT_MIN, bj = -1, bt = 0;
            for (int j = lane; j <= n; j += 32) {
                const int cc = j == 0 ? c0 : CC[j], dd = j == 0 ? c0 : DD[j];
                const int rr = j == n ? r0 : RR[n - j], ss = j == n ? r0 : SS[n - j];
                const int v1 = cc + rr, v2 = dd + ss - g;
                const int v = max(v1, v2);
                if (v > best) { best = v; bj = j; bt = !(v1 > v2); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(SEQA_FULL, best, o);
                const int oj = __shfl_xor_sync(SEQA_FULL, bj, o);
                const int ot = __shfl_xor_sync(SEQA_FULL, bt, o);
                if (oj >= 0 && (bj < 0 || ov > best || (ov == best && oj < bj))) { best = ov; bj = oj; bt = ot; }
            }
            if (lane == 0) {
                uint8_t *slot = A.slots + A.slot_off[p] + nd.i0 + nd.j0;
                LsNode l = nd, r = nd;
                l.q = 2 * nd.q;
                r.q = 2 * nd.q + 1;
                l.n = bj;
                r.j0 = nd.j0 + bj;
                r.n = n - bj;
                if (!bt) { // type 1 (:358-374)
                    l.m = mid; l.te = g;
                    r.i0 = nd.i0 + mid; r.m = m - mid; r.tb = g;
                } else { // type 2 (:375-395): rows mid-1 and mid are deleted
                    l.m = mid - 1; l.te = 0;
                    r.i0 = nd.i0 + mid + 1; r.m = m - mid - 1; r.tb = 0;
                    slot[(mid - 1) + bj] = 1;
                    slot[(mid - 1) + bj + 1] = 1;
                }
                ls_push(A, l);
                ls_push(A, r);
            }
        }
    }
}

// ---- squeeze the op slots + score the alignment ------------------------------------------------------------------
struct LsFinishArgs {
    const uint8_t *bases;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const uint32_t *idx;
    uint64_t count;
    uint8_t *slots;
    const uint64_t *slot_off;
    uint32_t *slot_start, *ops_len, *start_i, *start_j, *end_i, *end_j;
    int32_t *score;
    DevScoring sc;
    int affine;
};

// One warp per pair.  Score = the alignment re-scored under the algorithm's own gap model (the reference exposes
// no score for these two aligners; include/seqa_cuda.h documents the definition).
__global__ void __launch_bounds__(128) ls_finish_kernel(LsFinishArgs A)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t k = gw; k < A.count; k += nw) {
        const uint32_t p = A.idx[k];
        const int M = (int)A.len1[p], N = (int)A.len2[p];
        const uint8_t *a = A.bases + A.off1[p], *b = A.bases + A.off2[p];
        uint8_t *slot = A.slots + A.slot_off[p];
        int wr = 0, ci = 0, cj = 0, prev = -1;
        long long tot = 0;
        for (int base = 0; base < M + N; base += 32) {
            const int idx = base + lane;
            const unsigned op = idx < M + N ? slot[idx] : LS_HOLE;
            const bool live = op != LS_HOLE;
            const unsigned mlive = __ballot_sync(SEQA_FULL, live);
            const unsigned mi = __ballot_sync(SEQA_FULL, live && op != 2u);
            const unsigned mj = __ballot_sync(SEQA_FULL, live && op != 1u);
            // previous live op (for the affine run test)
            const unsigned below = mlive & lt;
            const int src = below ? 31 - __clz((int)below) : -1;
            const int pop = __shfl_sync(SEQA_FULL, (int)op, src < 0 ? 0 : src);
            const int pv = src < 0 ? prev : pop;
            if (live) {
                const int i = ci + __popc(mi & lt), j = cj + __popc(mj & lt);
                if (op == 0u)
                    tot += (a[i] == b[j]) ? A.sc.match : A.sc.mismatch;
                else if (A.affine)
                    tot += A.sc.ge + ((int)op != pv ? A.sc.go : 0);
                else
                    tot += A.sc.gap;
            }
            __syncwarp();
            if (live) slot[wr + __popc(below)] = (uint8_t)op;
            if (mlive) {
                const int last = 31 - __clz((int)mlive);
                prev = __shfl_sync(SEQA_FULL, (int)op, last);
            }
            wr += __popc(mlive);
            ci += __popc(mi);
            cj += __popc(mj);
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(SEQA_FULL, tot, o);
        if (lane == 0) {
            A.score[p] = (int32_t)tot;
            A.start_i[p] = 0;
            A.start_j[p] = 0;
            A.end_i[p] = (uint32_t)M;
            A.end_j[p] = (uint32_t)N;
            A.slot_start[p] = 0;
            A.ops_len[p] = (uint32_t)wr;
        }
    }
}

struct LsState {
    std::vector<LsNode> roots;
    std::vector<uint64_t> row_off;
    std::vector<uint32_t> row_w;
    uint64_t rows_total = 0;
    uint64_t node_cap = 0;
    bool mm = false;
    // device
    LsNode *d_nodes[2] = {nullptr, nullptr};
    uint32_t *d_count = nullptr; // cnt[4]
    int *d_overflow = nullptr;
    LsSweep *d_sweeps = nullptr;
    LsTask *d_tasks = nullptr;
    int *d_prog = nullptr;
    uint64_t task_cap = 0, sum_blocks = 0, cap_tasks = 0;
    int sweep_blocks = 0, sweep2_blocks = 0, sweep2_blocks_hb = 0;
    int *d_rows = nullptr;
    uint64_t *d_row_off = nullptr;
    uint32_t *d_row_w = nullptr;
    uint32_t *d_idx = nullptr;
    uint64_t cap_nodes = 0, cap_rows = 0, cap_pairs = 0;
    int levels_run = 0;
};
