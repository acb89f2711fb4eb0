// Linear-space global aligners on the GPU: HirschbergSA (reference include/SAHirschberg.h:11-184) and
// MyersMillerSA (reference include/SAMyersMiller.h:43-420).
//
// The reference recursion is run level by level: one launch per recursion depth processes every live
// sub-problem ("node") of every pair of the batch, one warp per node.  An internal node runs the forward and
// the reverse score-only sweep with the int32 warp-wavefront engine (seqa_wavefront.cuh, boundary values by
// __shfl_up_sync), finds the reference's split column with a warp reduction that reproduces its tie rule, and
// appends its two children to the next level's node list.  A leaf writes its ops straight into the pair's
// op slot at position i0 + j0 (at most i0 + j0 ops precede a node that starts at cell (i0, j0) and a node
// covering m rows and n columns emits at most m + n ops, so leaves never collide); a final pass squeezes out
// the unused slot bytes and scores the alignment.
//
// Bit-exactness notes: the recursion is followed all the way down to the
// reference's own leaves -- never short-circuited into a full-matrix aligner, because HirschbergSA's split
// search skips column N (include/SAHirschberg.h:141) and MyersMillerSA's M == 1 leaf is not the textbook one
// (include/SAMyersMiller.h:75-160): both are sub-optimal in a specific way that has to be reproduced.
#pragma once
#include <vector>
#include "seqa_common.cuh"
#include "seqa_wavefront.cuh"

#define LS_R 8
#define LS_HOLE 0xffu

struct LsArgs {
    const uint8_t *bases;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const LsNode *in;
    uint32_t n_in;
    LsNode *out;
    uint32_t *n_out;
    uint32_t out_cap;
    int *overflow;
    int *rows;               // scratch rows
    const uint64_t *row_off; // per pair: int offset of its 6 arrays
    const uint32_t *row_w;   // per pair: stride of one array
    uint8_t *slots;
    const uint64_t *slot_off;
    DevScoring sc;
};

__device__ __forceinline__ void ls_emit_run(uint8_t *slot, int from, int count, uint8_t op, int lane)
{
    for (int k = lane; k < count; k += 32) slot[from + k] = op;
}

__device__ __forceinline__ void ls_push(const LsArgs &A, const LsNode &nd)
{
    const uint32_t k = atomicAdd(A.n_out, 1u);
    if (k < A.out_cap)
        A.out[k] = nd;
    else
        *A.overflow = 1;
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

__global__ void __launch_bounds__(128) hb_level_kernel(LsArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    Borders bd;
    bd.hcolA = 0; bd.hcolB = A.sc.gap; bd.hrowA = 0; bd.hrowB = A.sc.gap;
    bd.ixA = bd.ixB = bd.iyA = bd.iyB = 0;
    for (uint32_t k = gw; k < A.n_in; k += nw) {
        const LsNode nd = A.in[k];
        const uint32_t p = (uint32_t)nd.pair;
        const uint8_t *a = A.bases + A.off1[p] + nd.i0;
        const uint8_t *b = A.bases + A.off2[p] + nd.j0;
        uint8_t *slot = A.slots + A.slot_off[p] + nd.i0 + nd.j0;
        const int m = nd.m, n = nd.n;
        if (m == 0) { // include/SAHirschberg.h:105
            ls_emit_run(slot, 0, n, 2, lane);
        } else if (n == 0) { // :112
            ls_emit_run(slot, 0, m, 1, lane);
        } else if (m == 1 || n == 1) { // :119
            hb_leaf_thin(A.sc, a, m, b, n, slot, lane);
        } else {
            const uint64_t w = A.row_w[p];
            int *base = A.rows + A.row_off[p] + nd.j0 + nd.q;
            int *F = base, *Rv = base + w, *BH = base + 2 * w;
            const int mid = m / 2; // :129
            wavefront<false, false, false, LS_R>(A.sc, bd, a, mid, b, n, false, nullptr, BH, nullptr, F, nullptr);
            __syncwarp();
            wavefront<false, false, false, LS_R>(A.sc, bd, a + mid, m - mid, b, n, true, nullptr, BH, nullptr, Rv, nullptr);
            if (lane == 0) {
                F[0] = mid * A.sc.gap;
                Rv[0] = (m - mid) * A.sc.gap;
            }
            __syncwarp();
            // Seq2Mid = argmax over i in [0, n-1] of F[i] + Rv[n-i], ties -> largest i (:141-149)
            int best = INT_MIN, bi = 0;
            for (int i = lane; i < n; i += 32) {
                const int v = F[i] + Rv[n - i];
                if (v >= best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(SEQA_FULL, best, o);
                const int oi = __shfl_xor_sync(SEQA_FULL, bi, o);
                if (ov > best || (ov == best && oi > bi)) { best = ov; bi = oi; }
            }
            __syncwarp();
            if (lane == 0) {
                LsNode l = nd, r = nd;
                l.m = mid; l.n = bi; l.q = 2 * nd.q;
                r.i0 = nd.i0 + mid; r.m = m - mid; r.j0 = nd.j0 + bi; r.n = n - bi; r.q = 2 * nd.q + 1;
                ls_push(A, l); // :151-155
                ls_push(A, r); // :157-161
            }
        }
    }
}

// ---- Myers-Miller ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mm_level_kernel(LsArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const int g = A.sc.go, h = A.sc.ge;
    for (uint32_t k = gw; k < A.n_in; k += nw) {
        const LsNode nd = A.in[k];
        const uint32_t p = (uint32_t)nd.pair;
        const uint8_t *a = A.bases + A.off1[p] + nd.i0;
        const uint8_t *b = A.bases + A.off2[p] + nd.j0;
        uint8_t *slot = A.slots + A.slot_off[p] + nd.i0 + nd.j0;
        const int M = nd.m, N = nd.n, tb = nd.tb, te = nd.te;
        if (N == 0) { // include/SAMyersMiller.h:57
            ls_emit_run(slot, 0, M, 1, lane);
        } else if (M == 0) { // :67
            ls_emit_run(slot, 0, N, 2, lane);
        } else if (M == 1) { // :75-160
            const int mx = max(tb, te) + h + (g + h * N);
            int best = INT_MIN, bj = 0;
            for (int j = 1 + lane; j <= N; j += 32) {
                const bool eq = a[0] == b[j - 1];
                int v = mx;
                if (A.sc.allow || eq) v = max(g + h * (j - 1) + (eq ? A.sc.match : A.sc.mismatch) + g + h * (N - j), mx); // :104-113
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
            ls_emit_run(slot, bj + (split ? 1 : 0), N - bj, 2, lane);
        } else {
            const uint64_t w = A.row_w[p];
            int *base = A.rows + A.row_off[p] + nd.j0 + nd.q;
            int *CC = base, *DD = base + w, *RR = base + 2 * w, *SS = base + 3 * w, *BH = base + 4 * w, *BX = base + 5 * w;
            const int mid = M / 2;
            Borders bf; // forward sweep borders (:172-198): H(i,0)=tb+i*h, H(0,j)=g+j*h, DD(0,j)=H(0,j)+g, e(i,0)=H(i,0)+g
            bf.hcolA = tb; bf.hcolB = h; bf.hrowA = g; bf.hrowB = h;
            bf.ixA = 2 * g; bf.ixB = h; bf.iyA = tb + g; bf.iyB = h;
            wavefront<true, false, false, LS_R>(A.sc, bf, a, mid, b, N, false, nullptr, BH, BX, CC, DD);
            __syncwarp();
            Borders br = bf; // reverse sweep (:247-313): same with te
            br.hcolA = te; br.iyA = te + g;
            wavefront<true, false, false, LS_R>(A.sc, br, a + mid, M - mid, b, N, true, nullptr, BH, BX, RR, SS);
            if (lane == 0) {
                CC[0] = tb + mid * h;
                DD[0] = CC[0]; // :238
                RR[0] = te + (M - mid) * h;
                SS[0] = RR[0]; // :313
            }
            __syncwarp();
            // midpoint (:315-340): first strictly greatest of max(CC+RR, DD+SS-g) over j = 0..N; RR/SS are stored
            // by reversed column index
            int best = INT_MIN, bj = -1, bt = 0;
            for (int j = lane; j <= N; j += 32) {
                const int v1 = CC[j] + RR[N - j], v2 = DD[j] + SS[N - j] - g;
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
            __syncwarp();
            if (lane == 0) {
                LsNode l = nd, r = nd;
                l.q = 2 * nd.q;
                r.q = 2 * nd.q + 1;
                l.n = bj;
                r.j0 = nd.j0 + bj;
                r.n = N - bj;
                if (!bt) { // type 1 (:358-374)
                    l.m = mid; l.te = g;
                    r.i0 = nd.i0 + mid; r.m = M - mid; r.tb = g;
                } else { // type 2 (:375-395): rows mid-1 and mid are deleted
                    l.m = mid - 1; l.te = 0;
                    r.i0 = nd.i0 + mid + 1; r.m = M - mid - 1; r.tb = 0;
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
    uint32_t *d_count = nullptr;
    int *d_overflow = nullptr;
    int *d_rows = nullptr;
    uint64_t *d_row_off = nullptr;
    uint32_t *d_row_w = nullptr;
    uint32_t *d_idx = nullptr;
    uint64_t cap_nodes = 0, cap_rows = 0, cap_pairs = 0;
    int levels_run = 0;
};
