// libseqa_cuda.so -- the C ABI of include/seqa_cuda.h: host-side batching / planning / sharding and the
// launches of the sm_100a kernels in seqa_packed.cuh (short linear-gap pairs, 2 pairs per thread, s16x2),
// seqa_wavefront.cuh (generic int32 warp wavefront, every full-matrix algorithm, any length) and
// seqa_linspace.cuh (Hirschberg / Myers-Miller recursion on the GPU).  No CPU fallback exists here: with no
// CUDA device every compute entry point fails with SEQA_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "seqa_common.cuh"
#include "seqa_util.cuh"
#include "seqa_wavefront.cuh"
#include "seqa_packed.cuh"
#include "seqa_packed_walk2.cuh"
#include "seqa_packed_affine.cuh"
#include "seqa_packed_affine_walk2.cuh"
#include "seqa_linspace.cuh"
#include "seqa_linspace_packed.cuh"
#include "../../include/seqa_cuda.h"

namespace {

thread_local std::string g_err;
thread_local std::vector<uint64_t> g_last_split; // cells per device of this thread's last seqa_cuda_align_batch call

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(SEQA_ERR_CUDA, "%s -> %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define CKS(call)                \
    do {                         \
        int s_ = (call);         \
        if (s_ != SEQA_OK) return s_; \
    } while (0)

template <class T> struct DBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return SEQA_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(SEQA_ERR_NOMEM, "device allocation of %zu bytes failed", n * sizeof(T));
        }
        cap = n;
        return SEQA_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// grow-only pinned host staging buffer: plan arrays are built here so that their upload is a true async DMA
// (a pageable source makes cudaMemcpyAsync wait behind every bulk copy queued on the copy engine)
template <class T> struct HBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return SEQA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(n + n / 4, 1024);
        if (cudaMallocHost((void **)&p, want * sizeof(T)) != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(SEQA_ERR_NOMEM, "pinned host allocation of %zu bytes failed", want * sizeof(T));
        }
        cap = want;
        return SEQA_OK;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Which packed fills rebuild the column profile from 2-bit column codes instead of reading 8 bytes per column, pair and
// strip (seqa_packed.cuh: pk_colprof): the affine fills, which are bound by HBM writes (+6-7 %: 2,601 -> 2,756 GCUPS
// GlobalGotoh, 2,616 -> 2,792 LocalGotoh).  The linear fills keep the precomputed profile: the short-pair ones are
// ALU-bound, and the long-pair ones measured slightly slower with codes.  DESIGN.md 4.
#ifndef PKG_CODES
#define PKG_CODES 1
#endif
#ifndef PK_GB_CODES_NW
#define PK_GB_CODES_NW 0 /* measured: 4,655 vs 4,752 GCUPS fill on the 50-1000 bp mix: the profile stays */
#endif
#ifndef PK_GB_CODES_SW
#define PK_GB_CODES_SW 0 /* 5,236 vs 5,375 */
#endif
constexpr int GEN_R = 4;          // rows per lane of the generic wavefront
constexpr int PK_R = 16;          // rows per register strip of the packed kernel
constexpr uint32_t PK_MAX_LEN = 320; // longest side the thread-per-pair kernel takes (shared-memory column)
constexpr uint32_t PKG_MAX_LEN = 2048; // affine thread-per-pair kernel: boundary rows live in global memory

struct Chunk {
    uint32_t lo, hi;       // range of jobs / list entries
    uint64_t scratch_bytes;
};

} // namespace

struct seqa_ctx;
namespace {
// linear-space algorithms (seqa_linspace_host.inl)
int ls_plan(seqa_ctx *c, const std::vector<uint32_t> &len1, const std::vector<uint32_t> &len2,
            const std::vector<uint32_t> &idx, bool myers_miller);
int ls_run(seqa_ctx *c, bool want_ops);
void ls_release(LsState &ls);
} // namespace

struct seqa_ctx {
    int device = 0;
    cudaStream_t stream = 0;    // kernels
    cudaStream_t up = 0, down = 0; // host->device / device->host copies; == stream except inside the pipelined one-shot call
    cudaEvent_t ev_up = nullptr, ev_run = nullptr; // order the three streams: uploads -> kernels -> downloads
    bool own_stream = false;
    int sms = 0;
    size_t smem_optin = 0;

    seqa_params prm{};
    DevScoring sc{};
    Borders bd{};
    uint64_t n = 0, cells = 0, slots_total = 0, bases_len = 0;
    std::vector<uint32_t> hlen1, hlen2;
    // batch statistics gathered by the one pass over the pairs that the upload makes anyway
    bool have_stats = false, st_uniform = false;
    uint64_t st_cells = 0, st_slots = 0;

    DBuf<uint8_t> bases;
    DBuf<uint8_t> packed_in;      // SEQA_FLAG_BASES_2BIT: the packed bytes as they crossed PCIe (unpacked into `bases`)
    DBuf<uint64_t> poff1, poff2;  // ... and their byte offsets when the batch is not dense
    DBuf<uint64_t> off1, off2;
    DBuf<uint32_t> len1, len2;

    DBuf<int32_t> score;
    DBuf<uint32_t> start_i, start_j, end_i, end_j, ops_len, slot_start;
    DBuf<uint64_t> slot_off, ops_off;
    DBuf<uint8_t> slots, dense;
    DBuf<uint64_t> tile_sum, total;
    DBuf<int> flags; // [0] packed path met a non-ACGT base, [1] job ticket of the packed fill kernels
    HBuf<uint64_t> h_tail; // pinned: [0] dense ops bytes of the last run, [1] flags[0]; copied at the end of every run

    // packed plan
    std::vector<uint32_t> pkl; // pairs the packed kernels take (kept between calls: no reallocation)
    HBuf<uint32_t> perm;
    size_t perm_n = 0;
    std::vector<PkWarpJob> jobs; // small: uploaded through jobs_pin
    HBuf<PkWarpJob> jobs_pin;
    size_t budget = 0;
    std::vector<Chunk> pk_chunks;
    uint32_t pk_max_nw = 0;
    DBuf<uint32_t> d_perm;
    DBuf<PkWarpJob> d_jobs;
    DBuf<uint4> pk_bound; // affine packed kernel: per-warp strip boundary rows
    // generic plan
    std::vector<uint32_t> gidx;
    std::vector<uint64_t> gdir_off;
    HBuf<uint32_t> gidx_pin;
    HBuf<uint64_t> gdir_pin;
    std::vector<Chunk> g_chunks;
    uint32_t g_max_n = 0;
    DBuf<uint32_t> d_gidx;
    DBuf<uint64_t> d_gdir_off;
    DBuf<int> bound;
    // linear-space plan
    std::vector<uint32_t> lidx;
    LsState ls;
    HBuf<uint64_t> ls_rowoff_pin;
    HBuf<uint32_t> ls_roww_pin, ls_idx_pin;

    // pairs rejected one by one (LocalGotoh shapes that are undefined behaviour in the reference)
    std::vector<uint32_t> ub_idx;
    HBuf<uint32_t> ub_pin;
    DBuf<uint32_t> d_ub;

    DBuf<uint8_t> scratch; // trace / direction matrices (+ profiles) of one chunk
    // pairs the packed path met a non-ACGT symbol in: flagged by pk_prep_kernel, re-run on the 8-bit kernels one by one
    // and kept off the packed path for later runs of the same resident batch
    DBuf<uint8_t> badpair;
    DBuf<uint32_t> d_badidx, d_badcnt;
    std::vector<uint8_t> forced_generic; // per pair, empty = none
    bool resolved = false;               // the last run's flagged pairs have been dealt with
    bool replan = false;                 // build_plan is re-planning a batch whose slots / result arrays exist already
    bool all_generic = false;            // linear-space path only: the batch holds non-ACGT symbols

    uint64_t ops_base = 0, ops_base_applied = 0; // offset added to ops_off by the next run / by the last run
    uint64_t launches = 0;
    cudaEvent_t dbg_ev[4] = {nullptr, nullptr, nullptr, nullptr}; // SEQA_DEBUG_TIMING: upload start / H2D done / kernels done / D2H done
    std::vector<cudaEvent_t> ev; // pairs of events around the DP fill launches of the last run
    size_t ev_used = 0;
    const char *last_kernel = "none";
    bool ran = false;
};

namespace {

#define LAUNCH(ctx, kern, grid, block, smem, ...)                               \
    do {                                                                        \
        (ctx)->launches++;                                                      \
        SEQA_LAUNCH(kern, grid, block, smem, (ctx)->stream, __VA_ARGS__);       \
    } while (0)

int validate_params(const seqa_params *p)
{
    if (!p) return fail(SEQA_ERR_INVALID, "params is NULL");
    if (p->algo < SEQA_NW || p->algo > SEQA_MYERS_MILLER) return fail(SEQA_ERR_INVALID, "bad algo %d", p->algo);
    const bool affine = p->algo == SEQA_GLOBAL_GOTOH || p->algo == SEQA_LOCAL_GOTOH || p->algo == SEQA_MYERS_MILLER;
    if (p->match <= 0) return fail(SEQA_ERR_UNSUPPORTED, "match must be > 0");
    if (p->allow_mismatch && p->mismatch >= 0) return fail(SEQA_ERR_UNSUPPORTED, "mismatch must be < 0");
    if (affine) {
        if (p->gap_open > 0 || p->gap_extend >= 0)
            return fail(SEQA_ERR_UNSUPPORTED, "need gap_open <= 0 and gap_extend < 0");
    } else if (p->gap >= 0)
        return fail(SEQA_ERR_UNSUPPORTED, "gap must be < 0");
    const int64_t lim = 1 << 20;
    if (std::abs((int64_t)p->match) > lim || std::abs((int64_t)p->gap) > lim || std::abs((int64_t)p->gap_open) > lim ||
        std::abs((int64_t)p->gap_extend) > lim || (p->allow_mismatch && std::abs((int64_t)p->mismatch) > lim))
        return fail(SEQA_ERR_UNSUPPORTED, "scoring magnitudes above 2^20 are not supported");
    return SEQA_OK;
}

void set_scoring(seqa_ctx *c)
{
    const seqa_params &p = c->prm;
    c->sc.gap = p.gap;
    c->sc.go = p.gap_open;
    c->sc.ge = p.gap_extend;
    c->sc.match = p.match;
    c->sc.mismatch = p.mismatch;
    c->sc.allow = p.allow_mismatch ? 1 : 0;
    Borders b{};
    switch (p.algo) {
    case SEQA_NW: // include/SANeedlemanWunsch.h:59-62
        b.hcolB = p.gap;
        b.hrowB = p.gap;
        break;
    case SEQA_GLOBAL_GOTOH: // include/SAGlobalGotoh.h:75-88
        b.hcolA = b.hrowA = p.gap_open;
        b.hcolB = b.hrowB = p.gap_extend;
        b.ixA = b.iyA = SEQA_GOTOH_NEG;
        break;
    case SEQA_LOCAL_GOTOH: // include/SALocalGotoh.h:77-90
        b.ixA = b.iyA = SEQA_GOTOH_NEG;
        break;
    default: break; // SW: zeros (include/SASmithWaterman.h:68-77); linear-space algorithms set their own
    }
    c->bd = b;
}

// Can the s16x2 thread-per-pair kernel take (M,N) under the current scoring?  See seqa_packed.cuh.
bool packed_affine(const seqa_params &p) { return p.algo == SEQA_GLOBAL_GOTOH || p.algo == SEQA_LOCAL_GOTOH; }
bool packed_scoring_ok(const seqa_params &p)
{
    if (p.flags & SEQA_FLAG_FORCE_GENERIC) return false;
    if (packed_affine(p)) {
        // seqa_packed_affine.cuh: profile scores minus (go+ge) must fit int8, every difference the walk tests on low
        // bytes must stay far inside (-128, 128), and the -128 "never" marker must lose against a gap open
        const int g = -(p.gap_open + p.gap_extend), m = p.match, x = p.allow_mismatch ? -p.mismatch : 0;
        if (m > 100 || g > 50 || x > 100 || m + g > 120) return false;
        if (m + x + 3 * g - p.gap_open > 120) return false;
        return true;
    }
    if (p.algo != SEQA_NW && p.algo != SEQA_SW) return false;
    const int g = -p.gap, m = p.match, x = p.allow_mismatch ? -p.mismatch : 0;
    if (m > 100 || g > 50 || x > 100) return false;
    if (m + x + 2 * g > 120) return false; // neighbouring cells must differ by < 128
    return true;
}
// Trace bits per cell of the packed linear path (seqa_packed.cuh): the walk's equality tests are exact modulo 2^TB when
// every difference it can meet lies in a window of at most 2^TB values that contains the tested constant.
//   8 bits: always (differences below 128: packed_scoring_ok);  4 bits: Match + |Mismatch| + 2|Gap| <= 7;
//   2 bits -- the algorithmic minimum -- for Match = 1, Gap = -1, Mismatch >= -2 (or no mismatches): the default
//   SmithWatermanSA scoring (-1, 1, -1).  Then H(i,j) - H(i-1,j-1) lies in [-2, 1] (tested against +1 / Mismatch),
//   H(i,j) - H(i-1,j) and H(i,j) - H(i,j-1) in [-1, 2] (tested against -1; the row scan decodes the window [-1, 2]);
//   NeedlemanWunsch's column-normalised K shifts the three windows by +1 / 0 / +1 and the constants with them.
int packed_trace_bits(const seqa_params &p)
{
    const int g = -p.gap, m = p.match, x = p.allow_mismatch ? -p.mismatch : 0;
    if (p.flags & SEQA_FLAG_TRACE8) return 8;
    if (m == 1 && g == 1 && x <= 2 && !(p.flags & SEQA_FLAG_TRACE4)) return 2;
    return m + x + 2 * g <= 7 ? 4 : 8;
}
// affine packed path: 4 trace bits per plane suffice when every difference the walk tests on low bits stays below 16
// (and the row scan of the local aligner inside [-8, 7]); bounds: seqa_packed_affine.cuh / DESIGN.md 4.3
int packed_affine_trace_bits(const seqa_params &p)
{
    const int g = -(p.gap_open + p.gap_extend), m = p.match, x = p.allow_mismatch ? -p.mismatch : 0;
    if (p.flags & SEQA_FLAG_TRACE8) return 8;
    return (m + x + 2 * g <= 12 && m + g <= 7) ? 4 : 8;
}
bool packed_shape_ok(const seqa_params &p, uint32_t M, uint32_t N)
{
    if (packed_affine(p)) {
        // 16-bit values, and nothing real may come near the reference's -10000 "minus infinity"
        if (M == 0 || N == 0 || M > PKG_MAX_LEN || N > PKG_MAX_LEN) return false;
        const int64_t unit = -(int64_t)(p.gap_open + p.gap_extend) + p.match + (p.allow_mismatch ? -p.mismatch : 0);
        return (int64_t)(M + N + 2 * PK_R + 2) * unit < 9000;
    }
    if (M == 0 || N == 0 || M > PKG_MAX_LEN || N > PKG_MAX_LEN) return false;
    const int64_t g = -p.gap, m = p.match;
    // NeedlemanWunsch keeps K = H - j*gap (seqa_packed.cuh): up to (N + padding) * |gap| above H
    const int64_t lo = (int64_t)(M + N + 2 * PK_R + 2) * g + 300;
    const int64_t hi = (int64_t)std::min(M, N) * m + (p.algo == SEQA_NW ? (int64_t)(N + 2 * PK_R + 2) * g : 0) + 300;
    return lo < 30000 && hi < 30000;
}

// scratch budget of one context (trace / direction matrices of one chunk): 80 % of the device shared by `sharers` + 1
// contexts -- a quarter of the device for a resident context, 0.8 / (ring + 1) for the contexts of a pipelined one-shot
// call (the producer lowers c->budget when it takes a context into its ring); queried once per context
size_t free_budget(int sharers = 3)
{
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return (size_t)1 << 30;
    return (size_t)std::min((double)fr * 0.8, (double)tot * 0.8 / (sharers + 1));
}
// the same cap from the device's total memory alone: cudaMemGetInfo costs ~1 ms, too much for the per-wave path of the
// one-shot call (the total is queried once per device and process)
size_t ring_budget(int device, int ring)
{
    static size_t total[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    const int d = device >= 0 && device < 64 ? device : 0;
    if (!total[d]) {
        size_t fr = 0, tot = 0;
        total[d] = cudaMemGetInfo(&fr, &tot) == cudaSuccess ? tot : ((size_t)8 << 30);
    }
    return (size_t)((double)total[d] * 0.8 / (ring + 1));
}

int order_after(seqa_ctx *c, cudaStream_t from, cudaStream_t to);

// generic jobs (one warp per pair) for the pairs listed in c->gidx: direction-matrix offsets, chunks, uploads
int plan_generic(seqa_ctx *c)
{
    const seqa_params &prm = c->prm;
    const size_t budget = c->budget;
    c->gdir_off.clear();
    c->g_chunks.clear();
    c->g_max_n = 0;
    if (!c->gidx.empty()) {
        const bool affine = prm.algo == SEQA_GLOBAL_GOTOH || prm.algo == SEQA_LOCAL_GOTOH;
        c->gdir_off.resize(c->gidx.size());
        Chunk ch{0, 0, 0};
        uint64_t words = 0;
        for (size_t k = 0; k < c->gidx.size(); k++) {
            const uint32_t p = c->gidx[k];
            const uint64_t wds = dir_words((int)c->hlen1[p], (int)c->hlen2[p], GEN_R, affine);
            if (wds * 4 > budget)
                return fail(SEQA_ERR_NOMEM, "pair %u (%u x %u) needs a %llu-byte direction matrix: use Hirschberg/MyersMiller",
                            p, c->hlen1[p], c->hlen2[p], (unsigned long long)(wds * 4));
            if (ch.hi > ch.lo && (words + wds) * 4 > budget) {
                ch.scratch_bytes = words * 4;
                c->g_chunks.push_back(ch);
                ch.lo = ch.hi;
                words = 0;
            }
            c->gdir_off[k] = words;
            words += wds;
            ch.hi = (uint32_t)(k + 1);
            c->g_max_n = std::max(c->g_max_n, c->hlen2[p]);
        }
        ch.scratch_bytes = words * 4;
        c->g_chunks.push_back(ch);
        CKS(c->d_gidx.ensure(c->gidx.size()));
        CKS(c->d_gdir_off.ensure(c->gidx.size()));
        CKS(c->gidx_pin.ensure(c->gidx.size()));
        CKS(c->gdir_pin.ensure(c->gidx.size()));
        std::copy(c->gidx.begin(), c->gidx.end(), c->gidx_pin.p);
        std::copy(c->gdir_off.begin(), c->gdir_off.end(), c->gdir_pin.p);
        CK(cudaMemcpyAsync(c->d_gidx.p, c->gidx_pin.p, c->gidx.size() * 4, cudaMemcpyHostToDevice, c->up));
        CK(cudaMemcpyAsync(c->d_gdir_off.p, c->gdir_pin.p, c->gidx.size() * 8, cudaMemcpyHostToDevice, c->up));
    }
    return SEQA_OK;
}

static int env_int(const char *name, int dflt, int lo, int hi);
static int pkg_ctas_per_sm();
static int pk_ctas_per_sm(size_t smem_optin, uint32_t cols);

int build_plan(seqa_ctx *c)
{
    const uint64_t n = c->n;
    const seqa_params &prm = c->prm;
    c->perm_n = 0;
    c->jobs.clear();
    c->pk_chunks.clear();
    c->gidx.clear();
    c->gdir_off.clear();
    c->g_chunks.clear();
    c->lidx.clear();
    c->pk_max_nw = 0;
    c->g_max_n = 0;
    if (!c->have_stats) {
        uint64_t cells = 0, run = 0;
        bool uni = true;
        for (uint64_t p = 0; p < n; p++) {
            cells += (uint64_t)c->hlen1[p] * c->hlen2[p];
            run += (uint64_t)c->hlen1[p] + c->hlen2[p];
            uni &= c->hlen1[p] == c->hlen1[0] && c->hlen2[p] == c->hlen2[0];
        }
        c->st_cells = cells;
        c->st_slots = run;
        c->st_uniform = uni;
        c->have_stats = true;
    }
    c->cells = c->st_cells;

    // per-pair op slots: len1+len2 bytes each (an alignment never has more columns); offsets by a device scan
    // (a re-plan of the same resident batch keeps them: ops_len holds results by then)
    if (!c->replan) {
        const uint64_t run = c->st_slots;
        c->slots_total = run;
        CKS(c->slot_off.ensure(n));
        CKS(c->slots.ensure(run + 32)); // gather_ops_kernel reads whole aligned (16-byte) words around a string
        CKS(c->ops_len.ensure(n));
        CKS(c->tile_sum.ensure((n + SEQA_SCAN_TILE - 1) / SEQA_SCAN_TILE + 1));
        CKS(c->total.ensure(1));
        if (n) {
            const unsigned tiles = (unsigned)((n + SEQA_SCAN_TILE - 1) / SEQA_SCAN_TILE);
            LAUNCH(c, (lensum_kernel), (unsigned)((n + 255) / 256), 256, 0, c->len1.p, c->len2.p, c->ops_len.p, n);
            LAUNCH(c, (scan_tile_sums_kernel), tiles, SEQA_SCAN_TPB, 0, c->ops_len.p, n, c->tile_sum.p, 0);
            LAUNCH(c, (scan_spine_kernel), 1, 1024, 0, c->tile_sum.p, (uint64_t)tiles, c->total.p);
            LAUNCH(c, (scan_apply_kernel), tiles, SEQA_SCAN_TPB, 0, c->ops_len.p, n, c->tile_sum.p, c->slot_off.p, 0, (uint64_t)0);
        }
    }
    CKS(c->score.ensure(n));
    CKS(c->start_i.ensure(n));
    CKS(c->start_j.ensure(n));
    CKS(c->end_i.ensure(n));
    CKS(c->end_j.ensure(n));
    CKS(c->ops_len.ensure(n));
    CKS(c->slot_start.ensure(n));
    CKS(c->ops_off.ensure(n));
    CKS(c->dense.ensure(c->slots_total));
    CKS(c->tile_sum.ensure((n + SEQA_SCAN_TILE - 1) / SEQA_SCAN_TILE + 1));
    CKS(c->total.ensure(1));
    CKS(c->flags.ensure(4));

    if (prm.algo == SEQA_HIRSCHBERG || prm.algo == SEQA_MYERS_MILLER) {
        c->lidx.resize(n);
        std::iota(c->lidx.begin(), c->lidx.end(), 0u);
        return ls_plan(c, c->hlen1, c->hlen2, c->lidx, prm.algo == SEQA_MYERS_MILLER);
    }
    // LocalGotoh shapes the reference routes into NW with an uninitialised Gap (SALocalGotoh.h:484-488, undefined
    // behaviour): rejected PER PAIR -- the pair is skipped and reported with ops_len = SEQA_PAIR_UNSUPPORTED, the rest
    // of the batch is aligned normally
    c->ub_idx.clear();
    auto ub_shape = [&](uint32_t M, uint32_t N) {
        return prm.algo == SEQA_LOCAL_GOTOH && ((M == 314 && N == 288) || (M == 60 && N == 57) || (M == 61 && N == 58));
    };
    if (prm.algo == SEQA_LOCAL_GOTOH) {
        if (c->st_uniform) {
            if (n && ub_shape(c->hlen1[0], c->hlen2[0])) {
                c->ub_idx.resize(n);
                std::iota(c->ub_idx.begin(), c->ub_idx.end(), 0u);
            }
        } else {
            for (uint64_t p = 0; p < n; p++)
                if (ub_shape(c->hlen1[p], c->hlen2[p])) c->ub_idx.push_back((uint32_t)p);
        }
        if (!c->ub_idx.empty()) {
            CKS(c->d_ub.ensure(c->ub_idx.size()));
            CKS(c->ub_pin.ensure(c->ub_idx.size()));
            std::copy(c->ub_idx.begin(), c->ub_idx.end(), c->ub_pin.p);
            CK(cudaMemcpyAsync(c->d_ub.p, c->ub_pin.p, c->ub_idx.size() * 4, cudaMemcpyHostToDevice, c->up));
        }
    }

    const bool pk = packed_scoring_ok(prm);
    const bool forced = !c->forced_generic.empty(); // pairs taken off the packed path (non-ACGT symbols seen by an earlier run)
    std::vector<uint32_t> &pkl = c->pkl;
    pkl.clear();
    bool uniform = true;
    // uniform batch (every pair the same shape): one eligibility test, identity permutation, identical jobs
    const bool fast = c->st_uniform && n > 0 && pk && packed_shape_ok(prm, c->hlen1[0], c->hlen2[0]) && c->ub_idx.empty() && !forced;
    if (fast) {
        pkl.resize(n);
        std::iota(pkl.begin(), pkl.end(), 0u);
    } else {
        if (pk) pkl.reserve(n);
        for (uint64_t p = 0; p < n; p++) {
            const uint32_t M = c->hlen1[p], N = c->hlen2[p];
            if (ub_shape(M, N)) continue;
            if (pk && packed_shape_ok(prm, M, N) && !(forced && c->forced_generic[p])) {
                if (!pkl.empty() && (M != c->hlen1[pkl[0]] || N != c->hlen2[pkl[0]])) uniform = false;
                pkl.push_back((uint32_t)p);
            } else {
                c->gidx.push_back((uint32_t)p);
            }
        }
    }
    if (!c->budget) c->budget = free_budget();
    if (const int kb = env_int("SEQA_SCRATCH_BUDGET_KB", 0, 0, 1 << 30)) c->budget = (size_t)kb << 10; // tests: many chunks from a small batch
    const size_t budget = c->budget;

    // ---- packed jobs: 64 pairs per warp, similar shapes together ----
    if (!pkl.empty()) {
        if (!uniform) {
            const std::vector<uint32_t> &l1 = c->hlen1, &l2 = c->hlen2;
            std::sort(pkl.begin(), pkl.end(), [&](uint32_t a, uint32_t b) {
                // largest first: the fill kernels hand jobs out through a ticket counter
                const uint32_t ka = (l1[a] + PK_R - 1) / PK_R, kb = (l1[b] + PK_R - 1) / PK_R;
                if (ka != kb) return ka > kb;
                if (l2[a] != l2[b]) return l2[a] > l2[b];
                return a < b;
            });
        }
        const size_t njobs = (pkl.size() + 63) / 64;
        CKS(c->perm.ensure(njobs * 64));
        c->perm_n = njobs * 64;
        std::copy(pkl.begin(), pkl.end(), c->perm.p);
        std::fill(c->perm.p + pkl.size(), c->perm.p + njobs * 64, PK_NULL);
        c->jobs.resize(njobs);
        Chunk ch{0, 0, 0};
        uint64_t tr = 0, pf = 0, rs = 0, lc = 0; // running offsets inside the chunk
        auto chunk_bytes = [&](uint64_t t, uint64_t q, uint64_t r, uint64_t l) { return t + q * 8 + r * 4 + l * 16 + 4096; };
        // A uniform batch larger than the scratch budget (10 M x 250 bp Gotoh: 98 KB of trace per pair) is cut at WHOLE ROUNDS of
        // the fill -- one job per resident warp: every warp then runs the same number of equal jobs and they end together.  Cut
        // at the budget alone (5,800 jobs = 3.27 rounds of 1,776) the last round ran with a quarter of the warps and took as
        // long as a full one.
        uint64_t cap_jobs = ~0ull;
        if (fast && c->sms > 0 && env_int("SEQA_ROUND_CHUNKS", 1, 0, 1)) {
            const uint32_t M = c->hlen1[0], N = c->hlen2[0], ns = (M + PK_R - 1) / PK_R;
            const uint64_t tb1 = packed_affine(prm) ? pkg_trace_bytes(ns, N, PK_R, packed_affine_trace_bits(prm)) : pk_trace_bytes(ns, N, PK_R, packed_trace_bits(prm));
            const uint64_t per_job = chunk_bytes(tb1, (uint64_t)((N + 3) / 4) * 128, pk_rowsel_elems(ns, N, PK_R), (uint64_t)ns * (PK_R / 4) * 32) - 4096;
            const uint64_t max_jobs = budget > 4096 ? (budget - 4096) / std::max<uint64_t>(per_job, 1) : 0;
            const uint64_t round = (uint64_t)c->sms * (uint64_t)(packed_affine(prm) ? pkg_ctas_per_sm() : pk_ctas_per_sm(c->smem_optin, N)) * (PK_BLOCK / 32);
            if (njobs > max_jobs && max_jobs >= round) cap_jobs = max_jobs / round * round;
        }
        for (size_t w = 0; w < njobs; w++) {
            uint32_t Mw = 0, Nw = 0;
            if (fast) {
                Mw = c->hlen1[0];
                Nw = c->hlen2[0];
            } else {
                for (int k = 0; k < 64; k++) {
                    const uint32_t p = c->perm.p[w * 64 + k];
                    if (p == PK_NULL) continue;
                    Mw = std::max(Mw, c->hlen1[p]);
                    Nw = std::max(Nw, c->hlen2[p]);
                }
            }
            PkWarpJob &J = c->jobs[w];
            J.first = (uint32_t)(w * 64);
            J.Mw = Mw;
            J.Nw = Nw;
            J.nstrips = (Mw + PK_R - 1) / PK_R;
            const uint64_t tbytes = packed_affine(prm) ? pkg_trace_bytes(J.nstrips, Nw, PK_R, packed_affine_trace_bits(prm))
                                                       : pk_trace_bytes(J.nstrips, Nw, PK_R, packed_trace_bits(prm));
            const uint64_t pelems = (uint64_t)((Nw + 3) / 4) * 128, relems = pk_rowsel_elems(J.nstrips, Nw, PK_R);
            const uint64_t lelems = (uint64_t)J.nstrips * (PK_R / 4) * 32; // last-column values (SW walk), uint4
            if (ch.hi > ch.lo && (chunk_bytes(tr + tbytes, pf + pelems, rs + relems, lc + lelems) > budget || ch.hi - ch.lo >= cap_jobs)) {
                ch.scratch_bytes = chunk_bytes(tr, pf, rs, lc);
                c->pk_chunks.push_back(ch);
                ch.lo = ch.hi;
                tr = pf = rs = lc = 0;
            }
            J.trace_off = tr;
            J.prof_off = pf;
            J.rowsel_off = rs;
            J.last_off = lc;
            tr += tbytes;
            pf += pelems;
            rs += relems;
            lc += lelems;
            ch.hi = (uint32_t)(w + 1);
            c->pk_max_nw = std::max(c->pk_max_nw, Nw);
        }
        ch.scratch_bytes = chunk_bytes(tr, pf, rs, lc);
        c->pk_chunks.push_back(ch);
        // chunk-relative layout: [trace | prof (8 B) | rowsel (4 B) | lastcol (16 B)], offsets resolved at launch
        CKS(c->d_perm.ensure(c->perm_n));
        CKS(c->d_jobs.ensure(c->jobs.size()));
        CKS(c->jobs_pin.ensure(c->jobs.size()));
        std::copy(c->jobs.begin(), c->jobs.end(), c->jobs_pin.p);
        CK(cudaMemcpyAsync(c->d_perm.p, c->perm.p, c->perm_n * 4, cudaMemcpyHostToDevice, c->up));
        CK(cudaMemcpyAsync(c->d_jobs.p, c->jobs_pin.p, c->jobs.size() * sizeof(PkWarpJob), cudaMemcpyHostToDevice, c->up));
    }

    CKS(plan_generic(c));
    uint64_t need = 16;
    for (auto &ch : c->pk_chunks) need = std::max(need, ch.scratch_bytes);
    for (auto &ch : c->g_chunks) need = std::max(need, ch.scratch_bytes);
    CKS(c->scratch.ensure(need));
    CKS(order_after(c, c->up, c->stream)); // the run reads perm / jobs
    return SEQA_OK; // everything above is stream-ordered; the host vectors it copies from are ctx members
}

// work queued on `to` from now on starts after everything queued on `from` so far (no-op for one stream)
int order_after(seqa_ctx *c, cudaStream_t from, cudaStream_t to)
{
    if (from == to) return SEQA_OK;
    cudaEvent_t e = from == c->stream ? c->ev_run : c->ev_up;
    CK(cudaEventRecord(e, from));
    CK(cudaStreamWaitEvent(to, e, 0));
    return SEQA_OK;
}

cudaEvent_t next_event(seqa_ctx *c)
{
    if (c->ev_used == c->ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev.push_back(e);
    }
    return c->ev[c->ev_used++];
}

template <bool AFFINE, bool LOCAL> void launch_generic(seqa_ctx *c, const Chunk &ch, int nwarps, int stride, bool want_walk)
{
    FillArgs F{};
    F.bases = c->bases.p;
    F.off1 = c->off1.p;
    F.off2 = c->off2.p;
    F.len1 = c->len1.p;
    F.len2 = c->len2.p;
    F.idx = c->d_gidx.p + ch.lo;
    F.count = ch.hi - ch.lo;
    F.dir_off = c->d_gdir_off.p + ch.lo;
    F.dir = reinterpret_cast<uint32_t *>(c->scratch.p);
    F.bound = c->bound.p;
    F.bound_stride = stride;
    F.score = c->score.p;
    F.end_i = c->end_i.p;
    F.end_j = c->end_j.p;
    F.sc = c->sc;
    F.bd = c->bd;
    const int blocks = (nwarps + 3) / 4;
    cudaEventRecord(next_event(c), c->stream);
    LAUNCH(c, (fill_i32_kernel<AFFINE, LOCAL, GEN_R>), blocks, 128, 0, F);
    cudaEventRecord(next_event(c), c->stream);
    if (!want_walk) return;
    WalkArgs W{};
    W.len1 = c->len1.p;
    W.len2 = c->len2.p;
    W.idx = F.idx;
    W.count = F.count;
    W.dir_off = F.dir_off;
    W.dir = F.dir;
    W.R = GEN_R;
    W.end_i = c->end_i.p;
    W.end_j = c->end_j.p;
    W.start_i = c->start_i.p;
    W.start_j = c->start_j.p;
    W.slots = c->slots.p;
    W.slot_off = c->slot_off.p;
    W.slot_start = c->slot_start.p;
    W.ops_len = c->ops_len.p;
    LAUNCH(c, (walk_kernel<AFFINE, LOCAL>), (unsigned)((F.count + 127) / 128), 128, 0, W);
}

int run_generic(seqa_ctx *c, bool want_walk)
{
    if (c->gidx.empty()) return SEQA_OK;
    const int stride = (int)((c->g_max_n + 1 + 63) / 32 * 32);
    for (const Chunk &ch : c->g_chunks) {
        const uint64_t cnt = ch.hi - ch.lo;
        int nwarps = (int)std::min<uint64_t>(cnt, (uint64_t)c->sms * 32);
        // keep the boundary scratch bounded for very long pairs
        while (nwarps > 4 && (uint64_t)nwarps * 2 * stride * 4 > ((uint64_t)1 << 30)) nwarps /= 2;
        nwarps = (nwarps + 3) / 4 * 4;
        CKS(c->bound.ensure((size_t)nwarps * 2 * stride));
        switch (c->prm.algo) {
        case SEQA_NW: launch_generic<false, false>(c, ch, nwarps, stride, want_walk); break;
        case SEQA_SW: launch_generic<false, true>(c, ch, nwarps, stride, want_walk); break;
        case SEQA_GLOBAL_GOTOH: launch_generic<true, false>(c, ch, nwarps, stride, want_walk); break;
        default: launch_generic<true, true>(c, ch, nwarps, stride, want_walk); break;
        }
        CK(cudaGetLastError());
    }
    if (c->last_kernel[0] == 'n') c->last_kernel = (c->prm.algo >= SEQA_GLOBAL_GOTOH) ? "fill_i32_affine" : "fill_i32_linear";
    return SEQA_OK;
}

static int env_int(const char *name, int dflt, int lo, int hi)
{
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    const long x = strtol(v, nullptr, 10);
    return (int)std::min<long>(hi, std::max<long>(lo, x));
}
// resident CTAs per SM of the packed affine fill (168 registers x 128 threads: 3 fit); SEQA_PKG_BPS overrides for A/B runs
// round-synchronous walk of the packed linear path (seqa_packed_walk2.cuh); SEQA_WALK2=0 selects pk_walk_kernel for A/B runs
static int use_walk2() { static const int v = PK_PAIR_PIECES != 0 ? 0 : env_int("SEQA_WALK2", PK_WALK2, 0, 2); return v; }
static int pkg_ctas_per_sm() { static const int v = env_int("SEQA_PKG_BPS", 3, 1, 3); return v; }

// resident CTAs per SM of the packed linear fill: its dynamic shared memory (one strip-boundary column per thread)
// is the limit, plus the 1 KB the driver reserves per CTA.  Used by the launch AND by the wave sizing of the one-shot
// call, which cuts waves at whole rounds of this kernel.
static int pk_ctas_per_sm(size_t smem_optin, uint32_t cols)
{
    if (cols > PK_MAX_LEN) return 3; // strip boundaries in global memory: registers are the limit
    const size_t smem = (size_t)cols * PK_BLOCK * 4;
    return (int)std::max<size_t>(1, std::min<size_t>(3, (smem_optin + 1024) / (smem + 1024)));
}

int run_packed(seqa_ctx *c, bool want_walk)
{
    if (c->jobs.empty()) return SEQA_OK;
    const bool affine = packed_affine(c->prm);
    const bool local = c->prm.algo == SEQA_SW || c->prm.algo == SEQA_LOCAL_GOTOH;
    const int tb = affine ? packed_affine_trace_bits(c->prm) : packed_trace_bits(c->prm);
    const bool gb = !affine && c->pk_max_nw > PK_MAX_LEN; // strip boundaries in global memory
    const size_t smem = (affine || gb) ? 0 : (size_t)c->pk_max_nw * PK_BLOCK * 4;
    if (smem > c->smem_optin) return fail(SEQA_ERR_UNSUPPORTED, "internal: packed kernel shared memory %zu", smem);
    if (smem) {
        CK(cudaFuncSetAttribute(pk_fill_kernel<true, PK_R, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(pk_fill_kernel<false, PK_R, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(pk_fill_kernel<true, PK_R, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(pk_fill_kernel<true, PK_R, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(pk_fill_kernel<false, PK_R, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(pk_fill_kernel<false, PK_R, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int bps = affine ? pkg_ctas_per_sm() : pk_ctas_per_sm(c->smem_optin, c->pk_max_nw);
    const uint64_t bound_stride = (uint64_t)((c->pk_max_nw + 3) / 4) * (affine ? 64 : 32);
    if (affine || gb) CKS(c->pk_bound.ensure((size_t)c->sms * bps * (PK_BLOCK / 32) * bound_stride));
    CK(cudaMemsetAsync(c->flags.p, 0, sizeof(int) * 4, c->stream));
    CKS(c->badpair.ensure(c->n));
    CK(cudaMemsetAsync(c->badpair.p, 0, c->n, c->stream));
    for (const Chunk &ch : c->pk_chunks) {
        const uint32_t nj = ch.hi - ch.lo;
        // chunk layout: trace | prof | rowsel | lastcol
        uint64_t tr = 0, pf = 0, rs = 0;
        {
            const PkWarpJob &L = c->jobs[ch.hi - 1];
            tr = L.trace_off + (affine ? pkg_trace_bytes(L.nstrips, L.Nw, PK_R, tb) : pk_trace_bytes(L.nstrips, L.Nw, PK_R, tb));
            pf = L.prof_off + (uint64_t)((L.Nw + 3) / 4) * 128;
            rs = L.rowsel_off + pk_rowsel_elems(L.nstrips, L.Nw, PK_R);
        }
        PkArgs A{};
        A.bases = c->bases.p;
        A.off1 = c->off1.p;
        A.off2 = c->off2.p;
        A.len1 = c->len1.p;
        A.len2 = c->len2.p;
        A.perm = c->d_perm.p;
        A.jobs = c->d_jobs.p + ch.lo;
        A.njobs = nj;
        A.trace = c->scratch.p;
        A.prof = reinterpret_cast<uint2 *>(c->scratch.p + ((tr + 255) / 256) * 256);
        A.rowsel = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(A.prof) + pf * 8);
        A.lastcol = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(A.rowsel) + ((rs * 4 + 15) / 16) * 16);
        A.score = c->score.p;
        A.end_i = c->end_i.p;
        A.end_j = c->end_j.p;
        A.start_i = c->start_i.p;
        A.start_j = c->start_j.p;
        A.slots = c->slots.p;
        A.slot_off = c->slot_off.p;
        A.slot_start = c->slot_start.p;
        A.ops_len = c->ops_len.p;
        A.bad = c->flags.p;
        A.badpair = c->badpair.p;
        A.gap = c->prm.gap;
        A.match = c->prm.match;
        A.mismatch = c->prm.mismatch;
        A.allow = c->prm.allow_mismatch ? 1 : 0;
        A.smem_cols = c->pk_max_nw;
        A.npos = (uint64_t)nj * 64;
        A.go = c->prm.gap_open;
        A.ge = c->prm.gap_extend;
        // subtracted from every profile score: affine G = H + go + ge; NeedlemanWunsch column-normalised (K = H - j*gap)
        A.prof_bias = affine ? c->prm.gap_open + c->prm.gap_extend : (local ? 0 : c->prm.gap);
        A.bound = c->pk_bound.p;
        A.bound_stride = bound_stride;
        A.ticket = reinterpret_cast<uint32_t *>(c->flags.p + 1);
        A.colcodes = affine ? (PKG_CODES != 0) : (gb && (local ? PK_GB_CODES_SW != 0 : PK_GB_CODES_NW != 0));
        CK(cudaMemsetAsync(c->flags.p + 1, 0, sizeof(int), c->stream));
        const unsigned wpb = PK_BLOCK / 32;
        const unsigned full = (nj + wpb - 1) / wpb;
        const unsigned grid = std::min<unsigned>(full, (unsigned)(c->sms * bps));
        // pk_prep_kernel: a uniform batch is dense in `bases` (both upload paths lay it out back to back), so the 128 sequences of
        // a job are one span of it -- staged in shared memory by one TMA bulk copy per job (two-warp CTAs: 5 per SM at 150 bp)
        unsigned prep_block = PK_BLOCK;
        size_t prep_smem = 0;
        A.prep_stage = 0;
        if (c->st_uniform && c->n > 0 && env_int("SEQA_PREP_TMA", 1, 0, 1)) {
            const uint64_t span = 64ull * ((uint64_t)c->hlen1[0] + c->hlen2[0]) + 32; // + 16-byte alignment at both ends
            if (span <= 48u * 1024u) {
                A.prep_stage = (uint32_t)((span + 15) & ~15ull);
                prep_block = 64;
                prep_smem = (size_t)(prep_block / 32) * (A.prep_stage + 64 + sizeof(uint64_t));
                if (prep_smem > 48u * 1024u)
                    CK(cudaFuncSetAttribute(pk_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
            }
        }
        const unsigned prep_wpb = prep_block / 32;
        LAUNCH(c, (pk_prep_kernel), (nj + prep_wpb - 1) / prep_wpb, prep_block, prep_smem, A, PK_R); // one job per warp
        cudaEventRecord(next_event(c), c->stream);
        if (affine && local && tb == 4)
            LAUNCH(c, (pkg_fill_kernel<true, PK_R, 4, PKG_CODES != 0>), grid, PK_BLOCK, 0, A);
        else if (affine && local)
            LAUNCH(c, (pkg_fill_kernel<true, PK_R, 8, PKG_CODES != 0>), grid, PK_BLOCK, 0, A);
        else if (affine && tb == 4)
            LAUNCH(c, (pkg_fill_kernel<false, PK_R, 4, PKG_CODES != 0>), grid, PK_BLOCK, 0, A);
        else if (affine)
            LAUNCH(c, (pkg_fill_kernel<false, PK_R, 8, PKG_CODES != 0>), grid, PK_BLOCK, 0, A);
        else if (gb && local && tb == 2)
            LAUNCH(c, (pk_fill_kernel<true, PK_R, 2, true, PK_GB_CODES_SW != 0>), grid, PK_BLOCK, 0, A);
        else if (gb && tb == 2)
            LAUNCH(c, (pk_fill_kernel<false, PK_R, 2, true, PK_GB_CODES_NW != 0>), grid, PK_BLOCK, 0, A);
        else if (local && tb == 2)
            LAUNCH(c, (pk_fill_kernel<true, PK_R, 2, false>), grid, PK_BLOCK, smem, A);
        else if (tb == 2)
            LAUNCH(c, (pk_fill_kernel<false, PK_R, 2, false>), grid, PK_BLOCK, smem, A);
        else if (gb && local && tb == 4)
            LAUNCH(c, (pk_fill_kernel<true, PK_R, 4, true, PK_GB_CODES_SW != 0>), grid, PK_BLOCK, 0, A);
        else if (gb && local)
            LAUNCH(c, (pk_fill_kernel<true, PK_R, 8, true, PK_GB_CODES_SW != 0>), grid, PK_BLOCK, 0, A);
        else if (gb && tb == 4)
            LAUNCH(c, (pk_fill_kernel<false, PK_R, 4, true, PK_GB_CODES_NW != 0>), grid, PK_BLOCK, 0, A);
        else if (gb)
            LAUNCH(c, (pk_fill_kernel<false, PK_R, 8, true, PK_GB_CODES_NW != 0>), grid, PK_BLOCK, 0, A);
        else if (local && tb == 4)
            LAUNCH(c, (pk_fill_kernel<true, PK_R, 4, false>), grid, PK_BLOCK, smem, A);
        else if (local)
            LAUNCH(c, (pk_fill_kernel<true, PK_R, 8, false>), grid, PK_BLOCK, smem, A);
        else if (tb == 4)
            LAUNCH(c, (pk_fill_kernel<false, PK_R, 4, false>), grid, PK_BLOCK, smem, A);
        else
            LAUNCH(c, (pk_fill_kernel<false, PK_R, 8, false>), grid, PK_BLOCK, smem, A);
        cudaEventRecord(next_event(c), c->stream);
        if (want_walk) {
            // walk positions are chunk-relative: perm/jobs pointers advanced to the chunk
            PkArgs Wk = A;
            Wk.perm = c->d_perm.p + (uint64_t)ch.lo * 64;
            // jobs' `first` fields are absolute; the walk indexes perm by position, so rebase via pointer only
            const unsigned wgrid = (unsigned)((Wk.npos + 255) / 256);
            const unsigned wgrid2 = (unsigned)((Wk.npos + PK_WALK2_TPB - 1) / PK_WALK2_TPB); // the round-synchronous walks
            // LocalGotoh stays on pkg_walk_kernel: its walk is the MaxCol row scan (125 pieces per 250 bp pair, at the HBM
            // roofline in both kernels: 0.57 ms per 200 k pairs there, 0.81 ms here), the alignments of the log regime are short
            if (affine && local && tb == 4 && use_walk2() > 1)
                LAUNCH(c, (pkg_walk2_kernel<true, PK_R>), wgrid2, PK_WALK2_TPB, 0, Wk);
            else if (affine && !local && tb == 4 && use_walk2())
                LAUNCH(c, (pkg_walk2_kernel<false, PK_R>), wgrid2, PK_WALK2_TPB, 0, Wk);
            else if (affine && local && tb == 4)
                LAUNCH(c, (pkg_walk_kernel<true, PK_R, 4>), wgrid, 256, 0, Wk);
            else if (affine && local)
                LAUNCH(c, (pkg_walk_kernel<true, PK_R, 8>), wgrid, 256, 0, Wk);
            else if (affine && tb == 4)
                LAUNCH(c, (pkg_walk_kernel<false, PK_R, 4>), wgrid, 256, 0, Wk);
            else if (affine)
                LAUNCH(c, (pkg_walk_kernel<false, PK_R, 8>), wgrid, 256, 0, Wk);
            else if (local && tb == 2 && use_walk2())
                LAUNCH(c, (pk_walk2_kernel<true, 2, PK_R>), wgrid2, PK_WALK2_TPB, 0, Wk);
            else if (tb == 2 && use_walk2())
                LAUNCH(c, (pk_walk2_kernel<false, 2, PK_R>), wgrid2, PK_WALK2_TPB, 0, Wk);
            else if (local && tb == 4 && use_walk2())
                LAUNCH(c, (pk_walk2_kernel<true, 4, PK_R>), wgrid2, PK_WALK2_TPB, 0, Wk);
            else if (tb == 4 && use_walk2())
                LAUNCH(c, (pk_walk2_kernel<false, 4, PK_R>), wgrid2, PK_WALK2_TPB, 0, Wk);
            else if (local && tb == 2)
                LAUNCH(c, (pk_walk_kernel<true, 2, PK_R>), wgrid, 256, 0, Wk);
            else if (tb == 2)
                LAUNCH(c, (pk_walk_kernel<false, 2, PK_R>), wgrid, 256, 0, Wk);
            else if (local && tb == 4)
                LAUNCH(c, (pk_walk_kernel<true, 4, PK_R>), wgrid, 256, 0, Wk);
            else if (local)
                LAUNCH(c, (pk_walk_kernel<true, 8, PK_R>), wgrid, 256, 0, Wk);
            else if (tb == 4)
                LAUNCH(c, (pk_walk_kernel<false, 4, PK_R>), wgrid, 256, 0, Wk);
            else
                LAUNCH(c, (pk_walk_kernel<false, 8, PK_R>), wgrid, 256, 0, Wk);
        }
        CK(cudaGetLastError());
    }
    if (affine)
        c->last_kernel = local ? (tb == 4 ? "pkg_fill_lgotoh_s16x2_t4" : "pkg_fill_lgotoh_s16x2_t8")
                               : (tb == 4 ? "pkg_fill_ggotoh_s16x2_t4" : "pkg_fill_ggotoh_s16x2_t8");
    else
        c->last_kernel = local ? (tb == 2 ? "pk_fill_sw_s16x2_t2" : tb == 4 ? "pk_fill_sw_s16x2_t4" : "pk_fill_sw_s16x2_t8")
                               : (tb == 2 ? "pk_fill_nw_s16x2_t2" : tb == 4 ? "pk_fill_nw_s16x2_t4" : "pk_fill_nw_s16x2_t8");
    return SEQA_OK;
}

int finish_ops(seqa_ctx *c)
{
    const uint64_t n = c->n;
    if (n == 0) return SEQA_OK;
    const unsigned tiles = (unsigned)((n + SEQA_SCAN_TILE - 1) / SEQA_SCAN_TILE);
    const int pack = (c->prm.flags & SEQA_FLAG_OPS_2BIT) ? 1 : 0; // dense ops: 4 per byte, offsets in bytes
    LAUNCH(c, (scan_tile_sums_kernel), tiles, SEQA_SCAN_TPB, 0, c->ops_len.p, n, c->tile_sum.p, pack);
    LAUNCH(c, (scan_spine_kernel), 1, 1024, 0, c->tile_sum.p, (uint64_t)tiles, c->total.p);
    // ops_off leaves the device already shifted by the wave's base inside the caller's ops buffer (one-shot pipeline): no
    // fix-up kernel on the download stream, where it would queue behind the next wave's persistent fill
    c->ops_base_applied = c->ops_base;
    LAUNCH(c, (scan_apply_kernel), tiles, SEQA_SCAN_TPB, 0, c->ops_len.p, n, c->tile_sum.p, c->ops_off.p, pack, c->ops_base);
    GatherArgs G{};
    G.n_pairs = n;
    G.slots = c->slots.p;
    G.slot_off = c->slot_off.p;
    G.slot_start = c->slot_start.p;
    G.ops_len = c->ops_len.p;
    G.ops_off = c->ops_off.p;
    G.dense = c->dense.p - c->ops_base; // dense index = ops_off - base
    G.pack = pack;
    // 8 lanes per short pair; few long pairs get enough warps for their bytes (the kernel then spreads a pair over warps)
    const unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((n * 8 + 255) / 256, c->slots_total / 16384 + 1), (uint64_t)c->sms * 64);
    LAUNCH(c, (gather_ops_kernel), blocks, 256, 0, G);
    CK(cudaGetLastError());
    return SEQA_OK;
}

int ctx_set_inputs_common(seqa_ctx *c, const seqa_params *params, uint64_t n)
{
    CKS(validate_params(params));
    c->prm = *params;
    c->n = n;
    c->ran = false;
    c->forced_generic.clear();
    c->replan = false;
    c->have_stats = false;
    c->ops_base = 0;
    set_scoring(c);
    CKS(c->off1.ensure(n));
    CKS(c->off2.ensure(n));
    CKS(c->len1.ensure(n));
    CKS(c->len2.ensure(n));
    return SEQA_OK;
}

// seqa_batch_in.sym_class -> the two 256-entry translation tables of translate_kernel: every class gets ONE representative
// byte (the classes of 'A', 'C', 'G', 'T' keep those letters, so that they stay on the packed 2-bit-code kernels; every
// other class a byte outside ACGT); the "matches nothing" class becomes 254 in sequence 1 and 255 in sequence 2, which
// never compare equal.  After this, byte equality IS the caller's match relation.
int build_class_tables(const uint8_t *cls, uint8_t *tab1, uint8_t *tab2)
{
    int rep[256];
    for (int &r : rep) r = -1;
    for (const char *l = "ACGT"; *l; l++) {
        const unsigned k = cls[(unsigned char)*l];
        if (k != SEQA_CLASS_NEVER && k < 254 && rep[k] < 0) rep[k] = (unsigned char)*l;
    }
    int next = 0;
    for (int ch = 0; ch < 256; ch++) {
        const unsigned k = cls[ch];
        if (k == SEQA_CLASS_NEVER) continue;
        if (k > 253) return fail(SEQA_ERR_INVALID, "sym_class[%d] = %u: class ids are 0..253 (255 = matches nothing)", ch, k);
        if (rep[k] < 0) {
            while (next == 'A' || next == 'C' || next == 'G' || next == 'T') next++;
            rep[k] = next++; // at most 250 such classes, 250 bytes outside {A,C,G,T,254,255}
        }
    }
    for (int ch = 0; ch < 256; ch++) {
        const unsigned k = cls[ch];
        tab1[ch] = k == SEQA_CLASS_NEVER ? (uint8_t)254 : (uint8_t)rep[k];
        tab2[ch] = k == SEQA_CLASS_NEVER ? (uint8_t)255 : (uint8_t)rep[k];
    }
    return SEQA_OK;
}

int translate_classes(seqa_ctx *c, const uint8_t *cls)
{
    if (!cls || c->n == 0) return SEQA_OK;
    TranslateArgs T{};
    CKS(build_class_tables(cls, T.tab1, T.tab2));
    T.bases = c->bases.p;
    T.off1 = c->off1.p;
    T.off2 = c->off2.p;
    T.len1 = c->len1.p;
    T.len2 = c->len2.p;
    T.n = c->n;
    const unsigned blocks = (unsigned)std::min<uint64_t>((2 * c->n * 32 + 255) / 256, (uint64_t)c->sms * 64);
    LAUNCH(c, (translate_kernel), blocks, 256, 0, T);
    CK(cudaGetLastError());
    return SEQA_OK;
}

// What one pass over a whole batch establishes, so that the waves of the one-shot call need no per-pair host loops of
// their own: every pair has the same shape, sequences lie back to back (seq1, seq2, next pair ...; byte-aligned sequences
// in the 2-bit wire format), and the last pair ends inside `bases`.
struct BatchFacts {
    bool uniform = false, dense = false, in_bounds = false;
};

// pairs [lo, hi) against the closed form anchored at pair 0 (the shape of pair 0, back to back from its offset on)
BatchFacts scan_batch_facts(const seqa_batch_in *in, bool two_bit, uint64_t lo0, uint64_t hi0, int max_threads = 8)
{
    BatchFacts f;
    const uint64_t n = in->n_pairs;
    if (n == 0 || hi0 <= lo0 || !in->off1 || !in->off2 || !in->len1 || !in->len2) return f;
    const uint32_t M0 = in->len1[0], N0 = in->len2[0];
    const uint64_t s1 = two_bit ? (uint64_t)((M0 + 3) >> 2) : (uint64_t)M0, s2 = two_bit ? (uint64_t)((N0 + 3) >> 2) : (uint64_t)N0;
    const uint64_t first = in->off1[0], stride = s1 + s2, cnt = hi0 - lo0;
    const int threads = (int)std::min<uint64_t>((uint64_t)max_threads, std::max<uint64_t>(1, cnt / 131072));
    std::vector<char> ok((size_t)threads, 1);
    auto part = [&](int t) {
        const uint64_t lo = lo0 + cnt * t / threads, hi = lo0 + cnt * (t + 1) / threads;
        unsigned diff = 0;
        uint64_t bad = 0;
        for (uint64_t p = lo; p < hi; p++) {
            diff |= (in->len1[p] ^ M0) | (in->len2[p] ^ N0);
            bad |= (in->off1[p] ^ (first + p * stride)) | (in->off2[p] ^ (first + p * stride + s1));
        }
        ok[(size_t)t] = (char)((diff == 0 ? 1 : 0) | (bad == 0 ? 2 : 0));
    };
    if (threads == 1) {
        part(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(part, t);
        part(0);
        for (auto &x : th) x.join();
    }
    f.uniform = f.dense = true;
    for (char v : ok) {
        f.uniform &= (v & 1) != 0;
        f.dense &= (v & 2) != 0;
    }
    f.dense &= f.uniform; // the closed form above only describes a uniform batch
    f.in_bounds = first <= in->bases_len && n * stride <= in->bases_len - first; // of the WHOLE batch, if it is uniform and dense
    return f;
}
BatchFacts scan_batch_facts(const seqa_batch_in *in, bool two_bit) { return scan_batch_facts(in, two_bit, 0, in->n_pairs); }

// SEQA_FLAG_BASES_2BIT: `in->bases` holds 2-bit symbols (4 per byte, A0 C1 T2 G3, every sequence on a byte boundary;
// off1 / off2 are BYTE offsets into it, len1 / len2 count symbols).  A quarter of the bytes cross PCIe; the device
// expands them once into the one-byte-per-symbol dense layout every kernel reads (unpack2_kernel).
int ctx_upload_range_2bit(seqa_ctx *c, const seqa_params *params, const seqa_batch_in *in, uint64_t pb, uint64_t pe, const BatchFacts *facts)
{
    const uint64_t n = pe - pb;
    CKS(ctx_set_inputs_common(c, params, n));
    c->hlen1.assign(in->len1 + pb, in->len1 + pe);
    c->hlen2.assign(in->len2 + pb, in->len2 + pe);
    uint64_t lo = UINT64_MAX, hi = 0, st_cells = 0, st_slots = 0, run_bytes = 0;
    bool st_uni = true, dense = true; // dense: seq1 then seq2 of every pair, byte-aligned, pairs back to back
    const uint32_t M0 = n ? in->len1[pb] : 0, N0 = n ? in->len2[pb] : 0;
    const uint64_t first_off = n ? in->off1[pb] : 0;
    const bool known = facts && facts->uniform && facts->dense && facts->in_bounds && n > 0; // established once for the whole batch
    if (known) {
        const uint64_t stride = (uint64_t)((M0 + 3) >> 2) + ((N0 + 3) >> 2);
        lo = first_off;
        hi = first_off + n * stride;
        st_cells = n * (uint64_t)M0 * N0;
        st_slots = n * ((uint64_t)M0 + N0);
    }
    for (uint64_t p = pb; p < pe && !known; p++) {
        const uint64_t l1 = in->len1[p], l2 = in->len2[p], B1 = (l1 + 3) >> 2, B2 = (l2 + 3) >> 2;
        const uint64_t a0 = in->off1[p], b0 = in->off2[p];
        if (a0 > in->bases_len || B1 > in->bases_len - a0 || b0 > in->bases_len || B2 > in->bases_len - b0)
            return fail(SEQA_ERR_INVALID, "pair %llu reaches past bases_len", (unsigned long long)p);
        dense &= a0 == first_off + run_bytes && b0 == a0 + B1;
        lo = std::min(lo, std::min(a0, b0));
        hi = std::max(hi, std::max(a0 + B1, b0 + B2));
        run_bytes += B1 + B2;
        st_cells += l1 * l2;
        st_slots += l1 + l2;
        st_uni &= l1 == M0 && l2 == N0;
    }
    if (getenv("SEQA_NO_DENSE_UPLOAD")) dense = false;
    c->have_stats = true;
    c->st_cells = st_cells;
    c->st_slots = st_slots;
    c->st_uniform = st_uni;
    if (n == 0 || hi < lo) lo = hi = 0;
    c->bases_len = st_slots; // the unpacked copy is dense whatever the layout of the packed bytes
    CKS(c->bases.ensure(c->bases_len + 64));
    CKS(c->packed_in.ensure(hi - lo + 16));
    if (hi > lo) CK(cudaMemcpyAsync(c->packed_in.p, in->bases + lo, hi - lo, cudaMemcpyHostToDevice, c->up));
    if (n) {
        if (!dense) {
            CKS(c->poff1.ensure(n));
            CKS(c->poff2.ensure(n));
            CK(cudaMemcpyAsync(c->poff1.p, in->off1 + pb, n * 8, cudaMemcpyHostToDevice, c->up));
            CK(cudaMemcpyAsync(c->poff2.p, in->off2 + pb, n * 8, cudaMemcpyHostToDevice, c->up));
        }
        if (!st_uni) {
            CK(cudaMemcpyAsync(c->len1.p, in->len1 + pb, n * 4, cudaMemcpyHostToDevice, c->up));
            CK(cudaMemcpyAsync(c->len2.p, in->len2 + pb, n * 4, cudaMemcpyHostToDevice, c->up));
        }
        CKS(order_after(c, c->up, c->stream));
        if (st_uni) LAUNCH(c, (fill_lengths_kernel), (unsigned)((n + 255) / 256), 256, 0, c->len1.p, c->len2.p, n, M0, N0);
        if (!dense && lo) LAUNCH(c, (rebase_kernel), (unsigned)((n + 255) / 256), 256, 0, c->poff1.p, c->poff2.p, n, lo);
    }
    int rc = build_plan(c);
    if (rc != SEQA_OK || n == 0) return rc;
    // unpacked offsets = the op-slot scan (exclusive scan of len1 + len2), as for a dense 8-bit batch
    LAUNCH(c, (dense_offsets_kernel), (unsigned)((n + 255) / 256), 256, 0, c->slot_off.p, c->len1.p, c->off1.p, c->off2.p, n);
    Unpack2Args U{};
    U.packed = c->packed_in.p;
    U.len1 = c->len1.p;
    U.len2 = c->len2.p;
    U.off1 = c->off1.p;
    U.off2 = c->off2.p;
    U.bases = c->bases.p;
    U.n = n;
    if (!dense) {
        U.poff1 = c->poff1.p;
        U.poff2 = c->poff2.p;
    } else if (st_uni && ((uint64_t)((M0 + 3) >> 2) + ((N0 + 3) >> 2)) > 0) {
        U.uniform_stride = (uint64_t)((M0 + 3) >> 2) + ((N0 + 3) >> 2);
    } else {
        // dense ragged batch: packed offsets = exclusive scan of ceil(len1/4) + ceil(len2/4), computed on the device
        CKS(c->poff1.ensure(n));
        const unsigned tiles = (unsigned)((n + SEQA_SCAN_TILE - 1) / SEQA_SCAN_TILE);
        LAUNCH(c, (packed_bytes_kernel), (unsigned)((n + 255) / 256), 256, 0, c->len1.p, c->len2.p, c->start_i.p, n); // start_i: scratch until the walk
        LAUNCH(c, (scan_tile_sums_kernel), tiles, SEQA_SCAN_TPB, 0, c->start_i.p, n, c->tile_sum.p, 0);
        LAUNCH(c, (scan_spine_kernel), 1, 1024, 0, c->tile_sum.p, (uint64_t)tiles, c->total.p);
        LAUNCH(c, (scan_apply_kernel), tiles, SEQA_SCAN_TPB, 0, c->start_i.p, n, c->tile_sum.p, c->poff1.p, 0, (uint64_t)0);
        U.poff1 = c->poff1.p;
    }
    const unsigned blocks = (unsigned)std::min<uint64_t>((2 * n * 32 + 255) / 256, (uint64_t)c->sms * 64);
    LAUNCH(c, (unpack2_kernel), blocks, 256, 0, U);
    CK(cudaGetLastError());
    return translate_classes(c, in->sym_class);
}

// upload pairs [pb, pe) of `in`; device offsets are rebased to the byte range the shard touches
int ctx_upload_range(seqa_ctx *c, const seqa_params *params, const seqa_batch_in *in, uint64_t pb, uint64_t pe, const BatchFacts *facts = nullptr)
{
    if (!in) return fail(SEQA_ERR_INVALID, "batch is NULL");
    const uint64_t n = pe - pb;
    if (n && (!in->bases || !in->off1 || !in->off2 || !in->len1 || !in->len2))
        return fail(SEQA_ERR_INVALID, "batch has NULL arrays");
    if (n > 0xfffffff0ull) return fail(SEQA_ERR_UNSUPPORTED, "more than 2^32-16 pairs per device shard");
    CK(cudaSetDevice(c->device));
    // Pipelined one-shot call: the planning kernels (length sums, scans, offsets, 2-bit unpack) go to the UPLOAD stream,
    // behind the copies they read, so that they run beside the previous wave's kernels instead of between two waves on
    // the compute stream; the compute stream waits once for everything (order_after below).
    struct PlanOnUploadStream {
        seqa_ctx *c;
        cudaStream_t saved;
        explicit PlanOnUploadStream(seqa_ctx *cx) : c(cx), saved(cx->stream) { c->stream = c->up; }
        ~PlanOnUploadStream() { c->stream = saved; }
    };
    if (c->up != c->stream) {
        int rc;
        {
            PlanOnUploadStream swap(c);
            rc = ctx_upload_range(c, params, in, pb, pe, facts);
        }
        if (rc == SEQA_OK) rc = order_after(c, c->up, c->stream);
        return rc;
    }
    if (params && (params->flags & SEQA_FLAG_BASES_2BIT)) return ctx_upload_range_2bit(c, params, in, pb, pe, facts);
    const bool dbgt = getenv("SEQA_DEBUG_TIMING") != nullptr;
    const auto tu0 = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    CKS(ctx_set_inputs_common(c, params, n));
    c->hlen1.assign(in->len1 + pb, in->len1 + pe);
    c->hlen2.assign(in->len2 + pb, in->len2 + pe);
    uint64_t lo = UINT64_MAX, hi = 0, st_cells = 0, st_slots = 0;
    bool st_uni = true, dense = true; // dense: seq1 then seq2 of every pair, pairs back to back from the first offset on
    const uint32_t M0 = n ? in->len1[pb] : 0, N0 = n ? in->len2[pb] : 0;
    const uint64_t first_off = n ? in->off1[pb] : 0;
    const bool known = facts && facts->uniform && facts->dense && facts->in_bounds && n > 0; // established once for the whole batch
    if (known) {
        lo = first_off;
        hi = first_off + n * ((uint64_t)M0 + N0);
        st_cells = n * (uint64_t)M0 * N0;
        st_slots = n * ((uint64_t)M0 + N0);
    }
    for (uint64_t p = pb; p < pe && !known; p++) {
        const uint64_t l1 = in->len1[p], l2 = in->len2[p];
        const uint64_t a0 = in->off1[p], a1 = a0 + l1, b0 = in->off2[p], b1 = b0 + l2;
        if (a0 > in->bases_len || l1 > in->bases_len - a0 || b0 > in->bases_len || l2 > in->bases_len - b0) // no wrap-around
            return fail(SEQA_ERR_INVALID, "pair %llu reaches past bases_len", (unsigned long long)p);
        dense &= a0 == first_off + st_slots && b0 == a1;
        lo = std::min(lo, std::min(a0, b0));
        hi = std::max(hi, std::max(a1, b1));
        st_cells += l1 * l2;
        st_slots += l1 + l2;
        st_uni &= l1 == M0 && l2 == N0;
    }
    if (getenv("SEQA_NO_DENSE_UPLOAD")) dense = false; // A/B switch: always send the offset / length arrays
    c->have_stats = true;
    c->st_cells = st_cells;
    c->st_slots = st_slots;
    c->st_uniform = st_uni;
    if (n == 0 || hi < lo) lo = hi = 0;
    c->bases_len = hi - lo;
    CKS(c->bases.ensure(c->bases_len + 64));
    if (hi > lo) CK(cudaMemcpyAsync(c->bases.p, in->bases + lo, hi - lo, cudaMemcpyHostToDevice, c->up));
    const bool dev_lengths = dense && st_uni; // uniform: the device fills the length arrays itself
    if (n) {
        // a dense batch sends no offsets (24 of its 324 bytes per 150 bp pair): the device derives them from the
        // op-slot scan of build_plan, which has the same values (exclusive scan of len1 + len2)
        if (!dense) {
            CK(cudaMemcpyAsync(c->off1.p, in->off1 + pb, n * 8, cudaMemcpyHostToDevice, c->up));
            CK(cudaMemcpyAsync(c->off2.p, in->off2 + pb, n * 8, cudaMemcpyHostToDevice, c->up));
        }
        if (!dev_lengths) {
            CK(cudaMemcpyAsync(c->len1.p, in->len1 + pb, n * 4, cudaMemcpyHostToDevice, c->up));
            CK(cudaMemcpyAsync(c->len2.p, in->len2 + pb, n * 4, cudaMemcpyHostToDevice, c->up));
        }
        CKS(order_after(c, c->up, c->stream)); // the planning kernels read len1/len2/off
        if (dev_lengths) LAUNCH(c, (fill_lengths_kernel), (unsigned)((n + 255) / 256), 256, 0, c->len1.p, c->len2.p, n, M0, N0);
        if (!dense && lo) LAUNCH(c, (rebase_kernel), (unsigned)((n + 255) / 256), 256, 0, c->off1.p, c->off2.p, n, lo);
    }
    const double t_up = ms_since(tu0);
    int rc = build_plan(c);
    if (rc == SEQA_OK && n && dense) // slot_off is ready (stream-ordered) and equals off1 relative to the shard's first byte
        LAUNCH(c, (dense_offsets_kernel), (unsigned)((n + 255) / 256), 256, 0, c->slot_off.p, c->len1.p, c->off1.p, c->off2.p, n);
    if (dbgt) fprintf(stderr, "[seqa]   upload %.3f ms, plan %.3f ms (%llu pairs)\n", t_up, ms_since(tu0) - t_up, (unsigned long long)n);
    if (rc == SEQA_OK) rc = translate_classes(c, in->sym_class); // off1 / off2 are final here (uploaded or derived)
    return rc;
}

// op strings gathered, per-pair status of rejected pairs set, downloads may follow
int finish_run(seqa_ctx *c)
{
    if (!c->ub_idx.empty()) // rejected pairs: no ops, neutral fields (the walk kernels never saw them)
        LAUNCH(c, (mark_pairs_kernel), (unsigned)((c->ub_idx.size() + 255) / 256), 256, 0, c->d_ub.p, (uint64_t)c->ub_idx.size(), c->score.p,
               c->start_i.p, c->start_j.p, c->end_i.p, c->end_j.p, c->ops_len.p, c->slot_start.p, 0u);
    CKS(finish_ops(c));
    if (!c->ub_idx.empty()) // ... and, once the op strings are gathered, the per-pair status the caller sees
        LAUNCH(c, (mark_pairs_kernel), (unsigned)((c->ub_idx.size() + 255) / 256), 256, 0, c->d_ub.p, (uint64_t)c->ub_idx.size(), c->score.p,
               c->start_i.p, c->start_j.p, c->end_i.p, c->end_j.p, c->ops_len.p, c->slot_start.p, (uint32_t)SEQA_PAIR_UNSUPPORTED);
    CK(cudaEventRecord(c->ev_run, c->stream)); // downloads are ordered behind this (ctx_resolve)
    return SEQA_OK;
}

int ctx_run(seqa_ctx *c)
{
    CK(cudaSetDevice(c->device));
    c->ev_used = 0;
    c->last_kernel = "none";
    if (c->n == 0) { c->ran = true; return SEQA_OK; }
    const bool want_walk = !(c->prm.flags & SEQA_FLAG_SCORE_ONLY);
    if (!want_walk) CK(cudaMemsetAsync(c->ops_len.p, 0, c->n * 4, c->stream));
    if (!c->lidx.empty()) {
        CKS(ls_run(c, want_walk));
    } else {
        CKS(run_packed(c, want_walk));
        CKS(run_generic(c, want_walk));
    }
    CKS(finish_run(c));
    c->ran = true;
    c->resolved = false;
    return SEQA_OK;
}

// The first synchronisation point after a run: fetch what the host must know (ctx_fetch_tail) and, if the packed path
// flagged pairs that hold a symbol outside ACGT, run exactly those pairs on the 8-bit kernels (ctx_resolve).
int ctx_fetch_tail(seqa_ctx *c)
{
    // what the host must know before it can fetch results (dense ops bytes, the packed path's "bad symbol" flag):
    // two small copies behind the run, one synchronisation.  Queued here, not at the end of ctx_run: on a download
    // stream shared by a ring of contexts they would sit in front of an earlier wave's result copies.
    CKS(c->h_tail.ensure(2));
    CK(cudaStreamWaitEvent(c->down, c->ev_run, 0));
    CK(cudaMemcpyAsync(&c->h_tail.p[0], c->total.p, 8, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(&c->h_tail.p[1], c->flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->down));
    CK(cudaStreamSynchronize(c->down));
    return SEQA_OK;
}

int ctx_resolve(seqa_ctx *c)
{
    if (!c->ran || c->n == 0) {
        CK(cudaStreamSynchronize(c->stream));
        return SEQA_OK;
    }
    CKS(ctx_fetch_tail(c));
    if (c->resolved || c->jobs.empty()) return SEQA_OK;
    c->resolved = true;
    const int bad = (int)(c->h_tail.p[1] & 0xffffffffu);
    if (!bad) return SEQA_OK;
    // Pairs with a symbol outside ACGT: their packed results are void.  Collect them, run THEM on the 8-bit kernels (the
    // other pairs keep their results), gather the op strings again, and keep them off the packed path for later runs.
    const uint64_t n = c->n;
    CKS(c->d_badidx.ensure(n));
    CKS(c->d_badcnt.ensure(1));
    CK(cudaMemsetAsync(c->d_badcnt.p, 0, 4, c->stream));
    LAUNCH(c, (collect_flagged_kernel), (unsigned)((n + 255) / 256), 256, 0, c->badpair.p, n, c->d_badidx.p, c->d_badcnt.p);
    uint32_t cnt = 0;
    CK(cudaMemcpyAsync(&cnt, c->d_badcnt.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> flagged(cnt);
    if (cnt) {
        CK(cudaMemcpyAsync(flagged.data(), c->d_badidx.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    std::sort(flagged.begin(), flagged.end());
    if (c->forced_generic.empty()) c->forced_generic.assign(n, 0);
    for (uint32_t p : flagged) c->forced_generic[p] = 1;
    c->gidx = flagged; // temporary plan: only the flagged pairs
    CKS(plan_generic(c));
    uint64_t need = 16;
    for (auto &ch : c->g_chunks) need = std::max(need, ch.scratch_bytes);
    CKS(c->scratch.ensure(std::max<uint64_t>(need, c->scratch.cap)));
    CKS(order_after(c, c->up, c->stream));
    const char *kernel = c->last_kernel;
    CKS(run_generic(c, !(c->prm.flags & SEQA_FLAG_SCORE_ONLY)));
    c->last_kernel = kernel; // the batch's dominant kernel stays the packed one
    CKS(finish_run(c));
    CKS(ctx_fetch_tail(c));
    c->replan = true; // later runs of this resident batch: the flagged pairs are planned onto the 8-bit kernels from the start
    const int rc = build_plan(c);
    c->replan = false;
    return rc;
}

int ctx_download_into(seqa_ctx *c, seqa_batch_out *out, uint64_t pb, uint64_t ops_base, uint64_t *ops_used)
{
    CK(cudaSetDevice(c->device));
    if (!c->ran) return fail(SEQA_ERR_INVALID, "download before run");
    CKS(ctx_resolve(c));
    const uint64_t n = c->n;
    *ops_used = 0;
    if (n == 0) return SEQA_OK;
    if (!out || !out->score) return fail(SEQA_ERR_INVALID, "out->score is NULL");
    CK(cudaMemcpyAsync(out->score + pb, c->score.p, n * 4, cudaMemcpyDeviceToHost, c->down));
    if (out->end_i) CK(cudaMemcpyAsync(out->end_i + pb, c->end_i.p, n * 4, cudaMemcpyDeviceToHost, c->down));
    if (out->end_j) CK(cudaMemcpyAsync(out->end_j + pb, c->end_j.p, n * 4, cudaMemcpyDeviceToHost, c->down));
    if (c->prm.flags & SEQA_FLAG_SCORE_ONLY) {
        CK(cudaStreamSynchronize(c->down));
        return SEQA_OK;
    }
    if (!out->start_i || !out->start_j || !out->end_i || !out->end_j || !out->ops || !out->ops_off || !out->ops_len)
        return fail(SEQA_ERR_INVALID, "output arrays are NULL (only allowed with SEQA_FLAG_SCORE_ONLY)");
    const uint64_t total = c->h_tail.p[0]; // valid: ctx_resolve synchronised the stream behind the run
    if (ops_base + total > out->ops_capacity)
        return fail(SEQA_ERR_CAPACITY, "ops_capacity %llu < %llu needed", (unsigned long long)out->ops_capacity,
                    (unsigned long long)(ops_base + total));
    const uint64_t delta = ops_base - c->ops_base_applied; // 0 when the run already placed ops_off (one-shot pipeline)
    if (delta) { c->launches++; SEQA_LAUNCH((add_base_kernel), (unsigned)((n + 255) / 256), 256, 0, c->down, c->ops_off.p, n, delta); }
    CK(cudaMemcpyAsync(out->start_i + pb, c->start_i.p, n * 4, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(out->start_j + pb, c->start_j.p, n * 4, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(out->ops_len + pb, c->ops_len.p, n * 4, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(out->ops_off + pb, c->ops_off.p, n * 8, cudaMemcpyDeviceToHost, c->down));
    if (total) CK(cudaMemcpyAsync(out->ops + ops_base, c->dense.p, total, cudaMemcpyDeviceToHost, c->down));
    if (delta) { c->launches++; SEQA_LAUNCH((add_base_kernel), (unsigned)((n + 255) / 256), 256, 0, c->down, c->ops_off.p, n, (uint64_t)0 - delta); } // restore: a second download stays valid
    CK(cudaStreamSynchronize(c->down));
    *ops_used = total;
    return SEQA_OK;
}

} // namespace

#include "seqa_linspace_host.inl"

extern "C" {

const char *seqa_cuda_last_error(void) { return g_err.c_str(); }
int seqa_cuda_abi_version(void) { return SEQA_ABI_VERSION; }

int seqa_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int seqa_ctx_create(seqa_ctx **out, int device, void *stream)
{
    if (!out) return fail(SEQA_ERR_INVALID, "ctx out pointer is NULL");
    *out = nullptr;
    const int nd = seqa_cuda_device_count();
    if (nd <= 0) return fail(SEQA_ERR_NO_DEVICE, "no CUDA device is visible (there is no CPU fallback)");
    if (device < 0 || device >= nd) return fail(SEQA_ERR_INVALID, "device %d out of range (0..%d)", device, nd - 1);
    CK(cudaSetDevice(device));
    seqa_ctx *c = new seqa_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c;
        return fail(SEQA_ERR_CUDA, "cudaGetDeviceProperties failed");
    }
#ifndef SEQA_EMU
    if (prop.major != 10) {
        delete c;
        return fail(SEQA_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
#endif
    c->sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return fail(SEQA_ERR_CUDA, "cudaStreamCreate failed");
        }
        c->own_stream = true;
    }
    c->up = c->down = c->stream;
    if (cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_run, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return fail(SEQA_ERR_CUDA, "cudaEventCreate failed");
    }
    *out = c;
    return SEQA_OK;
}

void seqa_ctx_destroy(seqa_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->bases.release(); c->packed_in.release(); c->poff1.release(); c->poff2.release(); c->off1.release(); c->off2.release(); c->len1.release(); c->len2.release();
    c->score.release(); c->start_i.release(); c->start_j.release(); c->end_i.release(); c->end_j.release();
    c->ops_len.release(); c->slot_start.release(); c->slot_off.release(); c->ops_off.release();
    c->slots.release(); c->dense.release(); c->tile_sum.release(); c->total.release(); c->flags.release();
    c->perm.release(); c->jobs_pin.release(); c->h_tail.release(); c->gidx_pin.release(); c->gdir_pin.release();
    c->ls_rowoff_pin.release(); c->ls_roww_pin.release(); c->ls_idx_pin.release();
    c->d_perm.release(); c->d_jobs.release(); c->pk_bound.release(); c->d_gidx.release(); c->d_gdir_off.release(); c->bound.release();
    c->scratch.release(); c->ub_pin.release(); c->d_ub.release(); c->badpair.release(); c->d_badidx.release(); c->d_badcnt.release();
    ls_release(c->ls);
    for (auto e : c->ev) cudaEventDestroy(e);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->ev_run) cudaEventDestroy(c->ev_run);
    for (auto e : c->dbg_ev)
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int seqa_ctx_upload(seqa_ctx *c, const seqa_params *params, const seqa_batch_in *in)
{
    if (!c || !in) return fail(SEQA_ERR_INVALID, "NULL argument");
    return ctx_upload_range(c, params, in, 0, in->n_pairs);
}

int seqa_ctx_generate(seqa_ctx *c, const seqa_params *params, uint64_t seed, uint64_t first_pair, uint64_t n_pairs,
                      int32_t len_mode, uint32_t len1, uint32_t len2)
{
    if (!c) return fail(SEQA_ERR_INVALID, "NULL ctx");
    if (n_pairs > 0xfffffff0ull) return fail(SEQA_ERR_UNSUPPORTED, "more than 2^32-16 pairs per device shard");
    CK(cudaSetDevice(c->device));
    CKS(ctx_set_inputs_common(c, params, n_pairs));
    c->hlen1.resize(n_pairs);
    c->hlen2.resize(n_pairs);
    std::vector<uint64_t> o1(n_pairs), o2(n_pairs);
    uint64_t run = 0;
    for (uint64_t p = 0; p < n_pairs; p++) {
        const uint32_t a = len_mode ? synth_len(seed, first_pair + p, 0) : len1;
        const uint32_t b = len_mode ? synth_len(seed, first_pair + p, 1) : len2;
        c->hlen1[p] = a;
        c->hlen2[p] = b;
        o1[p] = run;
        o2[p] = run + a;
        run += (uint64_t)a + b;
    }
    c->bases_len = run;
    CKS(c->bases.ensure(run + 64));
    if (n_pairs) {
        CK(cudaMemcpyAsync(c->off1.p, o1.data(), n_pairs * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->off2.p, o2.data(), n_pairs * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->len1.p, c->hlen1.data(), n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->len2.p, c->hlen2.data(), n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
        GenArgs G{seed, first_pair, n_pairs, c->off1.p, c->off2.p, c->len1.p, c->len2.p, c->bases.p};
        const unsigned blocks = (unsigned)std::min<uint64_t>((n_pairs * 32 + 255) / 256, (uint64_t)c->sms * 32);
        LAUNCH(c, (generate_kernel), blocks, 256, 0, G);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(c->stream));
    return build_plan(c);
}

int seqa_ctx_run(seqa_ctx *c)
{
    if (!c) return fail(SEQA_ERR_INVALID, "NULL ctx");
    return ctx_run(c);
}

int seqa_ctx_sync(seqa_ctx *c)
{
    if (!c) return fail(SEQA_ERR_INVALID, "NULL ctx");
    CK(cudaSetDevice(c->device));
    return ctx_resolve(c);
}

int seqa_ctx_device_results(seqa_ctx *c, seqa_batch_out *dev)
{
    if (!c || !dev) return fail(SEQA_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    if (!c->ran) return fail(SEQA_ERR_INVALID, "device results before run");
    CKS(ctx_resolve(c));
    dev->score = c->score.p;
    dev->start_i = c->start_i.p;
    dev->start_j = c->start_j.p;
    dev->end_i = c->end_i.p;
    dev->end_j = c->end_j.p;
    dev->ops = c->dense.p;
    dev->ops_off = c->ops_off.p;
    dev->ops_len = c->ops_len.p;
    dev->ops_used = c->n ? c->h_tail.p[0] : 0;
    dev->ops_capacity = dev->ops_used;
    return SEQA_OK;
}

int seqa_ctx_download(seqa_ctx *c, seqa_batch_out *out)
{
    if (!c || !out) return fail(SEQA_ERR_INVALID, "NULL argument");
    uint64_t used = 0;
    CKS(ctx_download_into(c, out, 0, 0, &used));
    out->ops_used = used;
    return SEQA_OK;
}

// Results of pairs [first, first + count) of the last run: the same arrays seqa_ctx_download fills, for a slice of a
// resident batch (a consumer that samples or streams results; bench.py's per-rank spot checks).  ops_off is
// rebased so that the slice's first op string starts at out->ops[0].
int seqa_ctx_download_range(seqa_ctx *c, uint64_t first, uint64_t count, seqa_batch_out *out)
{
    if (!c || !out) return fail(SEQA_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    if (!c->ran) return fail(SEQA_ERR_INVALID, "download before run");
    if (first > c->n || count > c->n - first) return fail(SEQA_ERR_INVALID, "pair range [%llu,+%llu) outside the batch of %llu",
                                                          (unsigned long long)first, (unsigned long long)count, (unsigned long long)c->n);
    CKS(ctx_resolve(c));
    out->ops_used = 0;
    if (count == 0) return SEQA_OK;
    if (!out->score) return fail(SEQA_ERR_INVALID, "out->score is NULL");
    const bool score_only = (c->prm.flags & SEQA_FLAG_SCORE_ONLY) != 0;
    CK(cudaMemcpyAsync(out->score, c->score.p + first, count * 4, cudaMemcpyDeviceToHost, c->down));
    if (out->end_i) CK(cudaMemcpyAsync(out->end_i, c->end_i.p + first, count * 4, cudaMemcpyDeviceToHost, c->down));
    if (out->end_j) CK(cudaMemcpyAsync(out->end_j, c->end_j.p + first, count * 4, cudaMemcpyDeviceToHost, c->down));
    if (score_only) {
        CK(cudaStreamSynchronize(c->down));
        return SEQA_OK;
    }
    if (!out->start_i || !out->start_j || !out->end_i || !out->end_j || !out->ops || !out->ops_off || !out->ops_len)
        return fail(SEQA_ERR_INVALID, "output arrays are NULL (only allowed with SEQA_FLAG_SCORE_ONLY)");
    CK(cudaMemcpyAsync(out->start_i, c->start_i.p + first, count * 4, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(out->start_j, c->start_j.p + first, count * 4, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(out->ops_len, c->ops_len.p + first, count * 4, cudaMemcpyDeviceToHost, c->down));
    CK(cudaMemcpyAsync(out->ops_off, c->ops_off.p + first, count * 8, cudaMemcpyDeviceToHost, c->down));
    uint64_t end = c->h_tail.p[0]; // dense bytes of the whole batch (ctx_resolve fetched it)
    if (first + count < c->n) CK(cudaMemcpyAsync(&end, c->ops_off.p + first + count, 8, cudaMemcpyDeviceToHost, c->down));
    CK(cudaStreamSynchronize(c->down));
    if (first + count == c->n) end += c->ops_base_applied;
    const uint64_t base = out->ops_off[0], bytes = end - base;
    if (bytes > out->ops_capacity)
        return fail(SEQA_ERR_CAPACITY, "ops_capacity %llu < %llu needed", (unsigned long long)out->ops_capacity, (unsigned long long)bytes);
    if (bytes) CK(cudaMemcpyAsync(out->ops, c->dense.p + (base - c->ops_base_applied), bytes, cudaMemcpyDeviceToHost, c->down));
    for (uint64_t k = 0; k < count; k++) out->ops_off[k] -= base;
    CK(cudaStreamSynchronize(c->down));
    out->ops_used = bytes;
    return SEQA_OK;
}

uint64_t seqa_ctx_launch_count(const seqa_ctx *c) { return c ? c->launches : 0; }
uint64_t seqa_ctx_cells(const seqa_ctx *c) { return c ? c->cells : 0; }
const char *seqa_ctx_last_kernel(const seqa_ctx *c) { return c ? c->last_kernel : "none"; }

int seqa_ctx_last_fill_ms(seqa_ctx *c, float *ms, int32_t *n_launches)
{
    if (!c || !ms) return fail(SEQA_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    float tot = 0.f;
    for (size_t k = 0; k + 1 < c->ev_used; k += 2) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, c->ev[k], c->ev[k + 1]));
        tot += t;
    }
    *ms = tot;
    if (n_launches) *n_launches = (int32_t)(c->ev_used / 2);
    return SEQA_OK;
}

int seqa_ctx_download_inputs(seqa_ctx *c, char *bases, uint64_t bases_len, uint64_t *off1, uint64_t *off2,
                             uint32_t *len1, uint32_t *len2)
{
    if (!c) return fail(SEQA_ERR_INVALID, "NULL ctx");
    CK(cudaSetDevice(c->device));
    if (bases_len < c->bases_len) return fail(SEQA_ERR_CAPACITY, "bases buffer too small");
    if (c->bases_len && bases) CK(cudaMemcpyAsync(bases, c->bases.p, c->bases_len, cudaMemcpyDeviceToHost, c->stream));
    if (c->n) {
        if (off1) CK(cudaMemcpyAsync(off1, c->off1.p, c->n * 8, cudaMemcpyDeviceToHost, c->stream));
        if (off2) CK(cudaMemcpyAsync(off2, c->off2.p, c->n * 8, cudaMemcpyDeviceToHost, c->stream));
        if (len1) CK(cudaMemcpyAsync(len1, c->len1.p, c->n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (len2) CK(cudaMemcpyAsync(len2, c->len2.p, c->n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return SEQA_OK;
}

// Lazily created per-device contexts reused by seqa_cuda_align_batch (device buffers survive between calls;
// seqa_cuda_trim() frees them).  Up to SEQA_CACHE_SLOTS contexts per device so that the waves of one call can be
// in flight together; a device whose cached contexts are all busy gets a temporary one.
#define SEQA_CACHE_SLOTS 8
#define SEQA_WORKERS 2 /* default host threads per device; each double-buffers two contexts */
static std::mutex g_cache_mu;
static cudaStream_t g_pipe_stream[64][4]; // per device: upload / kernel (even waves) / download / kernel (odd waves) streams of the pipelined one-shot call
static seqa_ctx *g_cache[64][SEQA_CACHE_SLOTS];
static bool g_cache_busy[64][SEQA_CACHE_SLOTS];

static int cache_acquire(int device, seqa_ctx **out, int *cached)
{
    *cached = -1;
    if (device >= 0 && device < 64) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (int k = 0; k < SEQA_CACHE_SLOTS; k++)
            if (g_cache[device][k] && !g_cache_busy[device][k]) {
                g_cache_busy[device][k] = true;
                *out = g_cache[device][k];
                *cached = k;
                return SEQA_OK;
            }
        for (int k = 0; k < SEQA_CACHE_SLOTS; k++)
            if (!g_cache[device][k]) {
                int s = seqa_ctx_create(out, device, nullptr);
                if (s != SEQA_OK) return s;
                g_cache[device][k] = *out;
                g_cache_busy[device][k] = true;
                *cached = k;
                return SEQA_OK;
            }
    }
    return seqa_ctx_create(out, device, nullptr);
}

static void cache_release(seqa_ctx *c, int cached)
{
    if (!c) return;
    if (cached < 0) {
        seqa_ctx_destroy(c);
        return;
    }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_busy[c->device][cached] = false;
}

int seqa_cuda_last_split(uint64_t *cells_per_device, int32_t capacity)
{
    for (int32_t d = 0; d < capacity && d < (int32_t)g_last_split.size(); d++)
        if (cells_per_device) cells_per_device[d] = g_last_split[d];
    return (int)g_last_split.size();
}

void *seqa_cuda_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (seqa_cuda_device_count() <= 0) {
        fail(SEQA_ERR_NO_DEVICE, "no CUDA device is visible (there is no CPU fallback)");
        return nullptr;
    }
    if (cudaHostAlloc(&p, bytes ? (size_t)bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        fail(SEQA_ERR_NOMEM, "cudaHostAlloc(%llu bytes) failed", (unsigned long long)bytes);
        return nullptr;
    }
    return p;
}

void seqa_cuda_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

void seqa_cuda_trim(void)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (int d = 0; d < 64; d++)
        for (int k = 0; k < SEQA_CACHE_SLOTS; k++)
            if (g_cache[d][k] && !g_cache_busy[d][k]) {
                seqa_ctx_destroy(g_cache[d][k]);
                g_cache[d][k] = nullptr;
            }
}

// One-shot entry.  The batch is cut into contiguous per-device shards (balanced by sum len1*len2) and every
// shard into WAVES of ~3e9 cells; up to three worker threads per device, each with its own context and stream,
// take the waves round-robin, so the host->device copy of one wave, the kernels of another and the device->host
// copy of a third overlap.  A wave's ops land in the caller's buffer at the prefix sum of (len1+len2) of the
// pairs before it (the ABI lets ops_off point anywhere), which needs no ordering between waves; when
// ops_capacity is smaller than sum(len1+len2) the waves are packed densely in order instead.
static int align_batch_impl(const seqa_params *params, const seqa_batch_in *in, seqa_batch_out *out, seqa_fill_fn fill, void *fill_user);

int seqa_cuda_align_batch(const seqa_params *params, const seqa_batch_in *in, seqa_batch_out *out)
{
    return align_batch_impl(params, in, out, nullptr, nullptr);
}

int seqa_cuda_align_batch_lazy(const seqa_params *params, const seqa_batch_in *in, seqa_batch_out *out, seqa_fill_fn fill, void *user)
{
    return align_batch_impl(params, in, out, fill, user);
}

static int align_batch_impl(const seqa_params *params, const seqa_batch_in *in, seqa_batch_out *out, seqa_fill_fn fill, void *fill_user)
{
    CKS(validate_params(params));
    if (!in || !out) return fail(SEQA_ERR_INVALID, "NULL argument");
    const int nd_all = seqa_cuda_device_count();
    if (nd_all <= 0) return fail(SEQA_ERR_NO_DEVICE, "no CUDA device is visible (there is no CPU fallback)");
    const int first = params->device_first;
    int nd = params->device_count > 0 ? params->device_count : nd_all - first;
    if (first < 0 || first >= nd_all || nd <= 0 || first + nd > nd_all)
        return fail(SEQA_ERR_INVALID, "device range [%d,%d) outside the %d visible devices", first, first + nd, nd_all);
    const uint64_t n = in->n_pairs;
    out->ops_used = 0;
    if (n == 0) return SEQA_OK;
    if (!in->len1 || !in->len2) return fail(SEQA_ERR_INVALID, "batch has NULL arrays");
    if ((uint64_t)nd > n) nd = (int)n;

    // one (threaded) pass over the index arrays: uniform shape? dense layout? inside `bases`?  The waves then skip their own loops
    const auto t_entry = std::chrono::steady_clock::now();
    const bool dbg_call = getenv("SEQA_DEBUG_TIMING") != nullptr;
    // A large batch on one device does not wait for that pass (0.4 ms per 1 M pairs, in front of everything): the waves are
    // planned as if the batch were uniform and dense like its first pair, the first wave's pairs are checked (a tenth of the
    // pass), and while that wave uploads and runs a helper thread checks the rest.  No later wave starts before the helper
    // agrees; if it does not, the first wave's results stand and the remaining pairs go through a second, ordinary call.
    const bool two_bit_in = (params->flags & SEQA_FLAG_BASES_2BIT) != 0;
    bool spec = n >= (uint64_t)env_int("SEQA_SPEC_MIN_PAIRS", 524288, 1, 1 << 30) && nd == 1 && in->off1 && in->off2 && !getenv("SEQA_NO_SPECULATIVE_FACTS");
    BatchFacts facts;
    if (spec) {
        facts = scan_batch_facts(in, two_bit_in, 0, 1, 1); // closed-form bounds of the whole batch from pair 0
        spec = facts.in_bounds;
    }
    if (!spec) facts = scan_batch_facts(in, two_bit_in);
    const double t_facts = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count();
    // waves: contiguous, ~3e9 cells or 256k pairs each; linear-space pairs (huge sweeps) go one wave per 4e10 cells
    const bool linspace = params->algo == SEQA_HIRSCHBERG || params->algo == SEQA_MYERS_MILLER;
    const uint64_t wave_cells = linspace ? 40000000000ull : (uint64_t)env_int("SEQA_WAVE_MCELLS", 3000, 10, 100000) * 1000000ull;
    std::vector<uint64_t> wave_lo, wave_slots, wave_cellsum; // first pair; (len1+len2) before the wave; cells of the wave
    uint64_t tot = 0;
    auto plan_waves = [&]() {
        wave_lo.clear();
        wave_slots.clear();
        wave_cellsum.clear();
        tot = 0;
        uint64_t acc = 0, slots = 0, cnt = 0;
        std::vector<uint64_t> sched; // explicit wave sizes (uniform packed batch on one device), else empty
        // several devices: at least two waves per device even for small batches
        uint64_t maxcnt = nd > 1 ? std::min<uint64_t>(262144, std::max<uint64_t>(1, (n + 2 * nd - 1) / (2 * nd))) : 262144;
        const uint32_t *l1 = in->len1, *l2 = in->len2;
        const bool uni = facts.uniform; // every pair the same shape: the waves are plain arithmetic
        // Uniform batch on the packed kernels: cut waves at whole "rounds" of the fill kernel (one 64-pair job per
        // resident warp), otherwise every wave pays a mostly idle last round.
        if (uni && !linspace && n > 1 && packed_scoring_ok(*params) && packed_shape_ok(*params, l1[0], l2[0])) {
            static int sms_cache[64];
            static size_t smem_cache[64];
            const int dev = first < 64 ? first : 0;
            if (!sms_cache[dev]) {
                cudaDeviceProp prop;
                if (cudaGetDeviceProperties(&prop, first) == cudaSuccess) {
                    smem_cache[dev] = prop.sharedMemPerBlockOptin;
                    sms_cache[dev] = prop.multiProcessorCount;
                }
                (void)cudaGetLastError();
            }
            if (sms_cache[dev] > 0) {
                const uint64_t per_cta = PK_BLOCK / 32 * 64;
                const uint64_t cta_per_sm = packed_affine(*params) ? (uint64_t)pkg_ctas_per_sm() : (uint64_t)pk_ctas_per_sm(smem_cache[dev], l2[0]);
                const uint64_t round = (uint64_t)sms_cache[dev] * cta_per_sm * per_cta;
                const uint64_t round_cells = round * ((uint64_t)l1[0] * l2[0] + 1);
                const uint64_t rounds = std::max<uint64_t>(1, (wave_cells + round_cells / 2) / round_cells);
                // whole rounds per wave only while every device still gets at least two waves (upload / kernels /
                // download of consecutive waves overlap); else keep the per-device cap, cut down to whole rounds
                if (round * rounds * 2 * (uint64_t)nd <= n) maxcnt = std::min<uint64_t>(n, round * rounds);
                else if (nd > 1 && maxcnt > round) maxcnt = maxcnt / round * round;
                // One device, several rounds, no explicit SEQA_WAVE_MCELLS: a ONE-round first wave (its upload is all that
                // delays the first kernel), then waves of SEQA_WAVE_ROUNDS rounds -- per-wave costs shrink; the last wave's walk,
                // gather and download are the tail.  Default 3 with pk_walk2_kernel (7.5-7.7 ms per 1 M x 150 bp against 7.6-7.9
                // with 2 and 7.8-8.1 with 4; pk_walk_kernel's 6 CTAs per SM preferred 2: two rounds of the fill were exactly one
                // full wave of its CTAs).
                if (nd == 1 && !getenv("SEQA_WAVE_MCELLS") && n > 2 * round) {
                    const uint64_t mid = (uint64_t)env_int("SEQA_WAVE_ROUNDS", 3, 1, 64);
                    uint64_t left = n - std::min(n, round);
                    sched.push_back(std::min(n, round));
                    while (left > 0) {
                        uint64_t take = std::min(left, mid * round);
                        if (left - take > 0 && left - take < round / 2) take = left; // no sliver at the end
                        sched.push_back(take);
                        left -= take;
                    }
                    // lazy fill: the caller's packing, not the GPU, paces the pipeline, and the GPU work on the LAST wave is
                    // what remains once the packing is done: keep it to one round
                    if (fill && sched.back() > round + round / 2) {
                        sched.back() -= round;
                        sched.push_back(round);
                    }
                }
            }
        }
        wave_lo.push_back(0);
        wave_slots.push_back(0);
        if (uni) {
            const uint64_t cells1 = (uint64_t)l1[0] * l2[0] + 1, slots1 = (uint64_t)l1[0] + l2[0];
            uint64_t per = maxcnt < 262144 ? maxcnt : std::min<uint64_t>(maxcnt, std::max<uint64_t>(1, (wave_cells + cells1 - 1) / cells1));
            size_t k = 0;
            for (uint64_t p = 0; p < n; p += per) {
                if (!sched.empty()) per = sched[std::min(k++, sched.size() - 1)];
                const uint64_t q = std::min(n, p + per);
                if (p) {
                    wave_lo.push_back(p);
                    wave_slots.push_back(p * slots1);
                }
                wave_cellsum.push_back((q - p) * cells1);
            }
            tot = n * cells1;
            slots = n * slots1;
        } else {
            for (uint64_t p = 0; p < n; p++) {
                if (cnt > 0 && (acc >= wave_cells || cnt >= maxcnt)) {
                    wave_lo.push_back(p);
                    wave_slots.push_back(slots);
                    wave_cellsum.push_back(acc);
                    acc = 0;
                    cnt = 0;
                }
                const uint64_t cells = (uint64_t)l1[p] * l2[p] + 1;
                acc += cells;
                tot += cells;
                slots += (uint64_t)l1[p] + l2[p];
                cnt++;
            }
            wave_cellsum.push_back(acc);
        }
        wave_lo.push_back(n);
        wave_slots.push_back(slots);
    };
    plan_waves();
    std::thread spec_thread;       // checks the pairs behind the first wave
    BatchFacts spec_rest;          // its verdict
    std::mutex spec_mu;
    std::condition_variable spec_cv;
    bool spec_done = false, spec_failed = false;
    if (spec) {
        bool ok = wave_lo.size() >= 3; // a single wave: nothing to overlap
        const bool ops2_ = (params->flags & SEQA_FLAG_OPS_2BIT) != 0;
        // the pipelined (sparse) layout only: the wave-after-wave path for small caller buffers plans with the full pass
        ok = ok && (out->ops_capacity >= (ops2_ ? wave_slots.back() / 4 + n : wave_slots.back()) || (params->flags & SEQA_FLAG_SCORE_ONLY));
        if (ok) {
            const BatchFacts f0 = scan_batch_facts(in, two_bit_in, 0, wave_lo[1], 1);
            ok = f0.uniform && f0.dense;
        }
        if (!ok) { // the ordinary full pass, and the waves planned from what it finds
            spec = false;
            facts = scan_batch_facts(in, two_bit_in);
            plan_waves();
        }
    }
    if (spec)
        spec_thread = std::thread([&]() {
            const BatchFacts r = scan_batch_facts(in, two_bit_in, wave_lo[1], n);
            std::lock_guard<std::mutex> lk(spec_mu);
            spec_rest = r;
            spec_done = true;
            spec_cv.notify_all();
        });
    struct SpecJoin { // the helper never outlives the call
        std::thread &t;
        ~SpecJoin() { if (t.joinable()) t.join(); }
    } spec_join{spec_thread};
    const size_t nwaves = wave_lo.size() - 1;
    const uint64_t slots_total = wave_slots.back();
    // where a wave's ops start in the caller's buffer: behind the slots of all earlier pairs (one byte per op), or, in the
    // 2-bit wire format, behind ceil(len/4) <= len/4 + 1 bytes per earlier pair
    const bool ops2 = (params->flags & SEQA_FLAG_OPS_2BIT) != 0;
    auto wave_ops_base = [&](size_t w) { return ops2 ? wave_slots[w] / 4 + wave_lo[w] : wave_slots[w]; };
    const bool sparse = out->ops_capacity >= (ops2 ? slots_total / 4 + n : slots_total) || (params->flags & SEQA_FLAG_SCORE_ONLY);
    // waves -> devices: contiguous runs of waves with ~equal cells (SURVEY.md 8e static split)
    std::vector<size_t> dev_lo(nd + 1, nwaves);
    dev_lo[0] = 0;
    if (nd > 1) {
        uint64_t run = 0;
        int d = 1;
        for (size_t w = 0; w < nwaves && d < nd; w++) {
            run += wave_cellsum[w];
            if ((long double)run >= (long double)tot * d / nd) dev_lo[d++] = w + 1;
        }
    }
    g_last_split.assign(nd, 0);
    for (int d = 0; d < nd; d++)
        for (size_t w = dev_lo[d]; w < dev_lo[d + 1]; w++) g_last_split[d] += wave_cellsum[w];
    std::vector<int> wstatus(nwaves, SEQA_OK);
    std::vector<std::string> werr(nwaves);
    std::vector<uint64_t> wused(nwaves, 0);
    const bool dbg = getenv("SEQA_DEBUG_TIMING") != nullptr;
    // SEQA_DEBUG_TIMING: per wave five events (H2D queued / H2D + planning done / kernels may start / kernels done /
    // D2H done) against one base event per device, printed as a table when the call ends
    struct WaveEv { cudaEvent_t e[5] = {}; int dev = -1; };
    std::vector<WaveEv> wev(dbg ? nwaves : 0);
    std::vector<cudaEvent_t> dbg_base(dbg ? nd : 0, nullptr);
    const auto t_start = std::chrono::steady_clock::now();
    if (dbg) fprintf(stderr, "[seqa] host before the pipeline: batch facts %.3f ms, wave schedule ..%.3f ms\n", t_facts,
                     std::chrono::duration<double, std::milli>(t_start - t_entry).count());
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    // Per device a PRODUCER thread uploads, plans and launches wave after wave into a ring of contexts (each with its
    // own stream) and a CONSUMER thread, in the same order, waits for a wave and copies its results out: the host
    // planning of wave k+1, the kernels of wave k and the device->host copy of wave k-1 overlap, and the device
    // queue is fed in wave order.
    const int ring = env_int("SEQA_RING", 4, 2, SEQA_CACHE_SLOTS);
    // kernels of consecutive waves on ONE stream by default: with two, the next wave's persistent fill and this wave's
    // walk fight for the same SMs and both lose (measured: 7.6 ms of kernel time per 1 M x 150 bp pairs against 6.6 ms
    // back to back, profiles/r02_e2e_timelines.txt); SEQA_TWO_COMPUTE_STREAMS=1 restores the even/odd streams
    const bool one_comp = getenv("SEQA_TWO_COMPUTE_STREAMS") == nullptr;
    struct DevPipe {
        std::mutex mu;
        std::condition_variable cv;
        long produced = 0, consumed = 0, filled = 0; // waves launched / downloaded / filled by the caller's callback
        bool stop = false;
        seqa_ctx *c[SEQA_CACHE_SLOTS] = {};
        int cached[SEQA_CACHE_SLOTS];
        cudaStream_t saved[SEQA_CACHE_SLOTS] = {};
        cudaStream_t st[4] = {}; // shared by the ring: uploads and downloads run in wave order; kernels of even and
                                 // odd waves go to two streams so that one wave's kernels fill the tail of the previous one's
    };
    std::vector<DevPipe> pipes(nd);
    auto producer = [&](int d) {
        DevPipe &P = pipes[d];
        const long nw = (long)(dev_lo[d + 1] - dev_lo[d]);
        for (long k = 0; k < nw; k++) {
            const size_t w = dev_lo[d] + (size_t)k;
            const int slot = (int)(k % ring);
            {
                std::unique_lock<std::mutex> lk(P.mu);
                P.cv.wait(lk, [&] { return P.consumed > k - ring || P.stop; });
                if (P.stop) break;
            }
            if (spec && w > 0) { // the pairs behind the first wave must have been checked, and found as assumed
                std::unique_lock<std::mutex> lk(spec_mu);
                spec_cv.wait(lk, [&] { return spec_done; });
                if (!(spec_rest.uniform && spec_rest.dense)) {
                    spec_failed = true;
                    break;
                }
            }
            int s = SEQA_OK;
            const double t0 = since();
            if (!P.c[slot]) {
                s = cache_acquire(first + d, &P.c[slot], &P.cached[slot]);
                if (s == SEQA_OK) {
                    const size_t b = ring_budget(first + d, ring);
                    if (!P.c[slot]->budget || P.c[slot]->budget > b) P.c[slot]->budget = b;
                }
                if (s == SEQA_OK && P.st[0]) {
                    seqa_ctx *cx = P.c[slot];
                    P.saved[slot] = cx->stream;
                    cx->up = P.st[0];
                    cx->down = P.st[2];
                }
            }
            if (s == SEQA_OK && P.st[0]) P.c[slot]->stream = (k & 1) && !one_comp ? P.st[3] : P.st[1];
            if (dbg && s == SEQA_OK) {
                cudaSetDevice(first + d);
                for (auto &e : wev[w].e) cudaEventCreate(&e);
                wev[w].dev = d;
                if (!dbg_base[d]) {
                    cudaEventCreate(&dbg_base[d]);
                    cudaEventRecord(dbg_base[d], P.c[slot]->up);
                }
                cudaEventRecord(wev[w].e[0], P.c[slot]->up);
            }
            if (fill) { // the filler thread runs ahead of this one: wait until it has produced this wave's symbols
                std::unique_lock<std::mutex> lk(P.mu);
                P.cv.wait(lk, [&] { return P.filled > k || P.stop; });
                if (P.stop) break;
            }
            if (s == SEQA_OK) s = ctx_upload_range(P.c[slot], params, in, wave_lo[w], wave_lo[w + 1], &facts);
            if (dbg && s == SEQA_OK) {
                cudaEventRecord(wev[w].e[1], P.c[slot]->up);
                cudaEventRecord(wev[w].e[2], P.c[slot]->stream);
            }
            const double t1 = since();
            if (s == SEQA_OK) P.c[slot]->ops_base = wave_ops_base(w);
            if (s == SEQA_OK) s = ctx_run(P.c[slot]);
            if (dbg && s == SEQA_OK) cudaEventRecord(wev[w].e[3], P.c[slot]->stream);
            if (dbg) fprintf(stderr, "[seqa] dev %d wave %zu: upload+plan %.2f..%.2f launched ..%.2f ms\n", d, w, t0, t1, since());
            std::lock_guard<std::mutex> lk(P.mu);
            if (s != SEQA_OK) {
                wstatus[w] = s;
                werr[w] = g_err;
                P.stop = true;
                P.cv.notify_all();
                break;
            }
            P.produced = k + 1;
            P.cv.notify_all();
        }
        std::lock_guard<std::mutex> lk(P.mu);
        P.stop = true;
        P.cv.notify_all();
    };
    // seqa_cuda_align_batch_lazy: a third thread per device calls the caller's fill callback wave after wave, ahead of the
    // producer, so that packing wave k+1, uploading / launching wave k and downloading wave k-1 all overlap
    auto filler = [&](int d) {
        DevPipe &P = pipes[d];
        const long nw = (long)(dev_lo[d + 1] - dev_lo[d]);
        for (long k = 0; k < nw; k++) {
            const size_t w = dev_lo[d] + (size_t)k;
            {
                std::lock_guard<std::mutex> lk(P.mu);
                if (P.stop) break;
            }
            const int rc = fill(fill_user, wave_lo[w], wave_lo[w + 1] - wave_lo[w]);
            std::lock_guard<std::mutex> lk(P.mu);
            if (rc != 0) {
                wstatus[w] = fail(SEQA_ERR_INVALID, "the caller's fill callback refused pairs [%llu, %llu)", (unsigned long long)wave_lo[w],
                                  (unsigned long long)wave_lo[w + 1]);
                werr[w] = g_err;
                P.stop = true;
                P.cv.notify_all();
                break;
            }
            P.filled = k + 1;
            P.cv.notify_all();
        }
    };
    auto consumer = [&](int d) {
        DevPipe &P = pipes[d];
        for (long k = 0;; k++) {
            {
                std::unique_lock<std::mutex> lk(P.mu);
                P.cv.wait(lk, [&] { return P.produced > k || P.stop; });
                if (P.produced <= k) break;
            }
            const size_t w = dev_lo[d] + (size_t)k;
            const int slot = (int)(k % ring);
            const double t2 = since();
            int s = ctx_resolve(P.c[slot]);
            const double t3 = since();
            if (s == SEQA_OK) s = ctx_download_into(P.c[slot], out, wave_lo[w], wave_ops_base(w), &wused[w]);
            if (dbg) {
                fprintf(stderr, "[seqa] dev %d wave %zu: wait %.2f..%.2f download ..%.2f ms\n", d, w, t2, t3, since());
                if (s == SEQA_OK && wev[w].e[4]) cudaEventRecord(wev[w].e[4], P.c[slot]->down);
            }
            std::lock_guard<std::mutex> lk(P.mu);
            if (s != SEQA_OK) {
                wstatus[w] = s;
                werr[w] = g_err;
                P.stop = true;
            }
            P.consumed = k + 1;
            P.cv.notify_all();
            if (s != SEQA_OK) break;
        }
    };
    int rc = SEQA_OK;
    uint64_t used = 0;
    if (sparse) {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; d++) {
            for (int q = 0; q < SEQA_CACHE_SLOTS; q++) pipes[d].cached[q] = -1;
            if (dev_lo[d + 1] == dev_lo[d]) continue;
            if (first + d < 64 && !getenv("SEQA_ONE_STREAM")) {
                std::lock_guard<std::mutex> lk(g_cache_mu);
                bool ok = cudaSetDevice(first + d) == cudaSuccess;
                for (int q = 0; q < 4 && ok; q++)
                    if (!g_pipe_stream[first + d][q])
                        ok = cudaStreamCreateWithFlags(&g_pipe_stream[first + d][q], cudaStreamNonBlocking) == cudaSuccess;
                if (ok)
                    for (int q = 0; q < 4; q++) pipes[d].st[q] = g_pipe_stream[first + d][q];
                (void)cudaGetLastError();
            }
            if (fill) th.emplace_back(filler, d);
            th.emplace_back(producer, d);
            th.emplace_back(consumer, d);
        }
        for (auto &t : th) t.join();
        if (dbg) fprintf(stderr, "[seqa] threads joined at %.3f ms\n", since());
        for (int d = 0; d < nd; d++)
            for (int q = 0; q < SEQA_CACHE_SLOTS; q++)
                if (pipes[d].c[q]) {
                    seqa_ctx *cx = pipes[d].c[q];
                    cudaSetDevice(cx->device);
                    cudaStreamSynchronize(cx->up);
                    cudaStreamSynchronize(cx->stream);
                    cudaStreamSynchronize(cx->down);
                    if (pipes[d].st[0]) cx->stream = pipes[d].saved[q]; // back to its own single stream
                    cx->up = cx->down = cx->stream;
                    cache_release(cx, pipes[d].cached[q]);
                }
        for (size_t w = 0; w < nwaves; w++) {
            if (rc == SEQA_OK && wstatus[w] != SEQA_OK) {
                rc = wstatus[w];
                g_err = werr[w];
            }
            used += wused[w];
        }
        if (spec_thread.joinable()) spec_thread.join();
        if (spec_failed && rc == SEQA_OK) {
            // the pairs behind the first wave are not like pair 0: they go through an ordinary call of their own (own pass
            // over the index arrays, own waves); their ops follow the first wave's slots in the caller's buffer
            const uint64_t pb = wave_lo[1], base = wave_ops_base(1);
            seqa_batch_in in2 = *in;
            in2.n_pairs = n - pb;
            in2.off1 += pb;
            in2.off2 += pb;
            in2.len1 += pb;
            in2.len2 += pb;
            seqa_batch_out out2 = *out;
            out2.score += pb;
            if (out2.start_i) out2.start_i += pb;
            if (out2.start_j) out2.start_j += pb;
            if (out2.end_i) out2.end_i += pb;
            if (out2.end_j) out2.end_j += pb;
            if (out2.ops_off) out2.ops_off += pb;
            if (out2.ops_len) out2.ops_len += pb;
            if (out2.ops) out2.ops += base;
            out2.ops_capacity = out->ops_capacity >= base ? out->ops_capacity - base : 0;
            out2.ops_used = 0;
            struct Shim {
                seqa_fill_fn fn;
                void *user;
                uint64_t off;
                static int call(void *u, uint64_t first_pair, uint64_t cnt)
                {
                    Shim *s = static_cast<Shim *>(u);
                    return s->fn(s->user, first_pair + s->off, cnt);
                }
            } shim{fill, fill_user, pb};
            rc = align_batch_impl(params, &in2, &out2, fill ? &Shim::call : nullptr, fill ? &shim : nullptr);
            if (rc == SEQA_OK && out->ops_off && !(params->flags & SEQA_FLAG_SCORE_ONLY))
                for (uint64_t q = pb; q < n; q++) out->ops_off[q] += base;
            used += out2.ops_used;
        }
        if (dbg) { // every stream was synchronised above: the events are complete
            fprintf(stderr, "[seqa] device timeline (ms after the first upload was queued): wave pairs | H2D start..done | kernels start..done | D2H done\n");
            for (size_t w = 0; w < nwaves; w++) {
                if (wev[w].dev < 0 || !wev[w].e[4] || !dbg_base[wev[w].dev]) continue;
                float t[5] = {0, 0, 0, 0, 0};
                bool ok = true;
                for (int q = 0; q < 5; q++) ok &= cudaEventElapsedTime(&t[q], dbg_base[wev[w].dev], wev[w].e[q]) == cudaSuccess;
                (void)cudaGetLastError();
                if (ok)
                    fprintf(stderr, "[seqa]   dev %d wave %2zu %7llu | %6.2f .. %6.2f | %6.2f .. %6.2f | %6.2f\n", wev[w].dev, w,
                            (unsigned long long)(wave_lo[w + 1] - wave_lo[w]), t[0], t[1], t[2], t[3], t[4]);
            }
            for (auto &we : wev)
                for (auto e : we.e)
                    if (e) cudaEventDestroy(e);
            for (auto e : dbg_base)
                if (e) cudaEventDestroy(e);
        }
    } else {
        // small caller buffer: waves one after another, ops packed densely in pair order
        for (int d = 0; d < nd && rc == SEQA_OK; d++) {
            seqa_ctx *c = nullptr;
            int cached = -1;
            rc = cache_acquire(first + d, &c, &cached);
            for (size_t w = dev_lo[d]; w < dev_lo[d + 1] && rc == SEQA_OK; w++) {
                uint64_t u = 0;
                if (fill && fill(fill_user, wave_lo[w], wave_lo[w + 1] - wave_lo[w]) != 0)
                    rc = fail(SEQA_ERR_INVALID, "the caller's fill callback refused pairs [%llu, %llu)", (unsigned long long)wave_lo[w],
                              (unsigned long long)wave_lo[w + 1]);
                if (rc != SEQA_OK) break;
                rc = ctx_upload_range(c, params, in, wave_lo[w], wave_lo[w + 1], &facts);
                if (rc == SEQA_OK) rc = ctx_run(c);
                if (rc == SEQA_OK) rc = ctx_download_into(c, out, wave_lo[w], used, &u);
                used += u;
            }
            if (c) cache_release(c, cached);
        }
    }
    out->ops_used = used;
    if (dbg_call) fprintf(stderr, "[seqa] call returns %.3f ms after entry\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
    return rc;
}

int seqa_cuda_int_peak(int device, int which, double *lane_ops_per_clk_per_sm, double *sm_clock_mhz)
{
#ifdef SEQA_EMU
    (void)device; (void)which; (void)lane_ops_per_clk_per_sm; (void)sm_clock_mhz;
    return fail(SEQA_ERR_NO_DEVICE, "micro-benchmark needs a GPU");
#else
    if (which < 0 || which > 9 || !lane_ops_per_clk_per_sm) return fail(SEQA_ERR_INVALID, "bad argument");
    if (seqa_cuda_device_count() <= device) return fail(SEQA_ERR_NO_DEVICE, "no such device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    unsigned *out = nullptr;
    long long *cyc = nullptr;
    CK(cudaMalloc((void **)&out, (size_t)blocks * 256 * 4));
    CK(cudaMalloc((void **)&cyc, (size_t)blocks * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0, 0));
        switch (which) {
        case 0: int_peak_kernel<0><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 1: int_peak_kernel<1><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 2: int_peak_kernel<2><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 3: int_peak_kernel<3><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 4: int_peak_kernel<4><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 5: int_peak_kernel<5><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 6: int_peak_kernel<6><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 7: int_peak_kernel<7><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        case 8: int_peak_kernel<8><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        default: int_peak_kernel<9><<<blocks, 256>>>(out, iters, 12345u, cyc); break;
        }
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    std::vector<long long> h(blocks);
    CK(cudaMemcpy(h.data(), cyc, (size_t)blocks * 8, cudaMemcpyDeviceToHost));
    long long cmax = 0;
    for (auto v : h) cmax = std::max(cmax, v);
    const double per_thread = (double)iters * 4 * 8 * (which == 9 ? 5 : 1);
    const double lane_ops = per_thread * 256.0 * blocks;
    // cycles: every SM runs 8 blocks concurrently (2048 threads); elapsed SM cycles ~ longest block
    *lane_ops_per_clk_per_sm = lane_ops / prop.multiProcessorCount / (double)cmax;
    if (sm_clock_mhz) *sm_clock_mhz = (double)cmax / (best * 1e3);
    cudaFree(out);
    cudaFree(cyc);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SEQA_OK;
#endif
}

} // extern "C"
