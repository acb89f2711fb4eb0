// Small device utilities of libseqa_cuda.so: exclusive scan, op-slot gather, synthetic-input generator,
// integer-pipe micro-benchmark.
#pragma once
#include "seqa_common.cuh"

__global__ void add_base_kernel(uint64_t *__restrict__ v, uint64_t n, uint64_t base)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) v[k] += base;
}

// ---- exclusive scan of uint32 -> uint64 (three launches; n up to 2^32) --------------------------------
// `pack`: the scanned item is ceil(in/4) -- the bytes of an op string in the 2-bit wire format (SEQA_FLAG_OPS_2BIT)
#define SEQA_SCAN_TPB 256
#define SEQA_SCAN_IPT 8
#define SEQA_SCAN_TILE (SEQA_SCAN_TPB * SEQA_SCAN_IPT)

__global__ void __launch_bounds__(SEQA_SCAN_TPB) scan_tile_sums_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                      uint64_t *__restrict__ tile_sum, int pack)
{
    __shared__ uint64_t wsum[SEQA_SCAN_TPB / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SEQA_SCAN_TILE;
    uint64_t s = 0;
    for (int k = 0; k < SEQA_SCAN_IPT; k++) {
        const uint64_t idx = base + (uint64_t)k * SEQA_SCAN_TPB + threadIdx.x;
        if (idx < n) s += pack ? (in[idx] + 3u) >> 2 : in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(SEQA_FULL, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < SEQA_SCAN_TPB / 32; w++) t += wsum[w];
        tile_sum[blockIdx.x] = t;
    }
}

// single block: in-place exclusive scan of the tile sums; total -> *total
__global__ void __launch_bounds__(1024) scan_spine_kernel(uint64_t *__restrict__ tile_sum, uint64_t ntiles,
                                                           uint64_t *__restrict__ total)
{
    __shared__ uint64_t part[1024];
    const uint64_t per = (ntiles + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = min(ntiles, per * threadIdx.x), hi = min(ntiles, lo + per);
    uint64_t s = 0;
    for (uint64_t k = lo; k < hi; k++) s += tile_sum[k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (unsigned k = 0; k < blockDim.x; k++) {
            const uint64_t v = part[k];
            part[k] = run;
            run += v;
        }
        *total = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint64_t k = lo; k < hi; k++) {
        const uint64_t v = tile_sum[k];
        tile_sum[k] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(SEQA_SCAN_TPB) scan_apply_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                  const uint64_t *__restrict__ tile_off,
                                                                  uint64_t *__restrict__ out, int pack, uint64_t add)
{
    // thread t owns the IPT consecutive items [base + t*IPT, +IPT); `add` is added to every output (a wave's base offset
    // inside the caller's ops buffer)
    __shared__ uint64_t wsum[SEQA_SCAN_TPB / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SEQA_SCAN_TILE + (uint64_t)threadIdx.x * SEQA_SCAN_IPT;
    uint32_t v[SEQA_SCAN_IPT];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SEQA_SCAN_IPT; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        if (pack) v[k] = (v[k] + 3u) >> 2;
        s += v[k];
    }
    // inclusive warp scan of s
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(SEQA_FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint64_t woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    uint64_t run = tile_off[blockIdx.x] + woff + inc - s + add;
#pragma unroll
    for (int k = 0; k < SEQA_SCAN_IPT; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(256) lensum_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                     uint32_t *__restrict__ out, uint64_t n)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = a[k] + b[k];
}

// Pairs rejected one by one (seqa_cuda.h: SEQA_PAIR_UNSUPPORTED): neutral result fields and the given ops_len
// (0 before the op strings are gathered, the status marker after).
__global__ void __launch_bounds__(256) mark_pairs_kernel(const uint32_t *__restrict__ idx, uint64_t n, int32_t *__restrict__ score,
                                                         uint32_t *__restrict__ start_i, uint32_t *__restrict__ start_j,
                                                         uint32_t *__restrict__ end_i, uint32_t *__restrict__ end_j,
                                                         uint32_t *__restrict__ ops_len, uint32_t *__restrict__ slot_start, uint32_t len_value)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = idx[k];
    score[p] = INT32_MIN;
    start_i[p] = start_j[p] = end_i[p] = end_j[p] = 0u;
    slot_start[p] = 0u;
    ops_len[p] = len_value;
}

// Dense batches (seq1 then seq2 of every pair, pairs back to back: what packers produce) need no offset arrays over
// PCIe: off1 = exclusive scan of (len1 + len2) -- the same scan that places the op slots -- and off2 = off1 + len1.
__global__ void __launch_bounds__(256) dense_offsets_kernel(const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ len1,
                                                            uint64_t *__restrict__ off1, uint64_t *__restrict__ off2, uint64_t n)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const uint64_t o = slot_off[k];
        off1[k] = o;
        off2[k] = o + len1[k];
    }
}

// uniform batches do not need the length arrays either
__global__ void __launch_bounds__(256) fill_lengths_kernel(uint32_t *__restrict__ len1, uint32_t *__restrict__ len2, uint64_t n, uint32_t a, uint32_t b)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        len1[k] = a;
        len2[k] = b;
    }
}

// device offsets are relative to the first byte the shard touches
__global__ void __launch_bounds__(256) rebase_kernel(uint64_t *__restrict__ off1, uint64_t *__restrict__ off2, uint64_t n, uint64_t lo)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        off1[k] -= lo;
        off2[k] -= lo;
    }
}

// Pairs the packed path flagged (a symbol outside ACGT): their indices, in any order, behind a counter.
__global__ void __launch_bounds__(256) collect_flagged_kernel(const uint8_t *__restrict__ flag, uint64_t n, uint32_t *__restrict__ idx,
                                                              uint32_t *__restrict__ count)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && flag[k]) idx[atomicAdd(count, 1u)] = (uint32_t)k;
}

// seqa_batch_in.sym_class: every symbol is replaced by a representative of its class, once, so that every kernel of the
// path keeps comparing bytes with == .  tab1 / tab2 translate sequence 1 / sequence 2 (they differ only for the
// "matches nothing" class, which maps to two different bytes).  One warp per sequence.
struct TranslateArgs {
    uint8_t *bases;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    uint64_t n;
    uint8_t tab1[256], tab2[256];
};

__global__ void __launch_bounds__(256) translate_kernel(TranslateArgs A)
{
    __shared__ uint8_t t1[256], t2[256];
    t1[threadIdx.x] = A.tab1[threadIdx.x];
    t2[threadIdx.x] = A.tab2[threadIdx.x];
    __syncthreads();
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t q = gw; q < 2 * A.n; q += nw) {
        const uint64_t p = q >> 1;
        const bool second = (q & 1) != 0;
        const uint32_t len = second ? A.len2[p] : A.len1[p];
        uint8_t *s = A.bases + (second ? A.off2[p] : A.off1[p]);
        const uint8_t *t = second ? t2 : t1;
        for (uint32_t k = lane; k < len; k += 32) s[k] = t[s[k]];
    }
}

// ---- SEQA_FLAG_BASES_2BIT: 2-bit symbols over PCIe, unpacked on the device -------------------------------------------
// Wire format: 4 symbols per byte (symbol k of a sequence in bits 2*(k%4) of byte k/4), code A0 C1 T2 G3 = (letter >> 1) & 3,
// every sequence on a byte boundary.  The unpacked copy (one byte per symbol, seq1 then seq2 per pair, dense) is what
// every kernel of the path reads, exactly as if the caller had sent 8-bit symbols: a quarter of the PCIe bytes for one
// streaming pass over HBM (0.4 GB per 1 M x 150 bp pairs).
__global__ void __launch_bounds__(256) packed_bytes_kernel(const uint32_t *__restrict__ len1, const uint32_t *__restrict__ len2,
                                                           uint32_t *__restrict__ out, uint64_t n)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = ((len1[k] + 3u) >> 2) + ((len2[k] + 3u) >> 2);
}

struct Unpack2Args {
    const uint8_t *packed;           // the shard's packed bytes
    const uint64_t *poff1, *poff2;   // packed byte offsets; poff2 == nullptr: sequence 2 follows sequence 1 (dense)
    uint64_t uniform_stride;         // != 0: poff1[p] = p * uniform_stride (uniform dense batch, no offset array at all)
    const uint32_t *len1, *len2;
    const uint64_t *off1, *off2;     // unpacked offsets
    uint8_t *bases;
    uint64_t n;
};

__global__ void __launch_bounds__(256) unpack2_kernel(Unpack2Args A)
{
    // one warp per sequence; lane l expands packed bytes l, l + 32, ... (4 symbols each)
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t q = gw; q < 2 * A.n; q += nw) {
        const uint64_t p = q >> 1;
        const bool second = (q & 1) != 0;
        const uint32_t l1 = A.len1[p], len = second ? A.len2[p] : l1;
        uint64_t po = A.uniform_stride ? p * A.uniform_stride : A.poff1[p];
        if (second) po = A.poff2 ? A.poff2[p] : po + ((l1 + 3u) >> 2);
        uint8_t *dst = A.bases + (second ? A.off2[p] : A.off1[p]);
        const uint32_t nb = (len + 3u) >> 2;
        for (uint32_t b = lane; b < nb; b += 32) {
            const unsigned v = A.packed[po + b];
            // "ACTG"[code] for the four codes of the byte
            const unsigned w = ((0x47544341u >> (8u * (v & 3u))) & 0xffu) | (((0x47544341u >> (8u * ((v >> 2) & 3u))) & 0xffu) << 8) |
                               (((0x47544341u >> (8u * ((v >> 4) & 3u))) & 0xffu) << 16) | (((0x47544341u >> (8u * ((v >> 6) & 3u))) & 0xffu) << 24);
            uint8_t *d = dst + 4u * b;
            const uint32_t have = len - 4u * b;
            if (have >= 4u && (reinterpret_cast<uintptr_t>(d) & 3u) == 0) {
                *reinterpret_cast<uint32_t *>(d) = w;
            } else {
                for (uint32_t t = 0; t < 4u && t < have; t++) d[t] = (uint8_t)(w >> (8u * t));
            }
        }
    }
}

// ---- gather the per-pair op slots (written back-to-front by the walk kernels) into the dense buffer ----
struct GatherArgs {
    uint64_t n_pairs;
    const uint8_t *slots;
    const uint64_t *slot_off;
    const uint32_t *slot_start;
    const uint32_t *ops_len;
    const uint64_t *ops_off; // exclusive scan of ops_len
    uint8_t *dense;
    int pack; // SEQA_FLAG_OPS_2BIT: 4 ops per dense byte, ops_off in bytes
};

__global__ void __launch_bounds__(256) gather_ops_kernel(GatherArgs A)
{
    // Batches with fewer pairs than warps (the 100 kbp linear-space pairs): every pair is spread over `wpp` warps.
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (A.n_pairs < nw) {
        const uint32_t sub = threadIdx.x & 31;
        const uint64_t wpp = nw / A.n_pairs;
        const uint64_t p = gw / wpp;
        if (p >= A.n_pairs) return;
        const uint32_t first = (uint32_t)(gw % wpp) * 32u + sub, step = (uint32_t)wpp * 32u;
        const uint8_t *src = A.slots + A.slot_off[p] + A.slot_start[p];
        uint8_t *dst = A.dense + A.ops_off[p];
        const uint32_t len = A.ops_len[p];
        if (A.pack) {
            for (uint32_t b = first; (uint64_t)b * 4 < len; b += step) {
                unsigned v = 0;
#pragma unroll
                for (uint32_t q = 0; q < 4; q++)
                    if (b * 4 + q < len) v |= (unsigned)(src[b * 4 + q] & 3u) << (2 * q);
                dst[b] = (uint8_t)v;
            }
        } else {
            for (uint32_t k = first; k < len; k += step) dst[k] = src[k];
        }
        return;
    }
    // Many short pairs: 8 lanes per pair.  The op strings sit at arbitrary byte offsets on both sides; copied byte by byte the
    // kernel was bound by its load / store instructions (a warp instruction moved 32 bytes).  It now works on the 32-bit words of
    // the DESTINATION grid: every interior word is built from aligned source words by a funnel shift (2-bit output: from two
    // aligned 16-byte loads, 16 ops -> one word) and leaves as one 32-bit store; only the first and the last word of a string,
    // which are shared with the neighbouring strings, go byte by byte.  Aligned words that hold a byte of the string lie inside
    // the buffers (256-byte aligned allocations, 32 bytes of slack behind `slots`).
    const uint32_t sub = threadIdx.x & 7;
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const uint64_t ng = ((uint64_t)gridDim.x * blockDim.x) >> 3;
    for (uint64_t p = g; p < A.n_pairs; p += ng) {
        const uint8_t *src = A.slots + A.slot_off[p] + A.slot_start[p];
        uint8_t *dst = A.dense + A.ops_off[p];
        const uint32_t len = A.ops_len[p];
        const uint32_t nb = A.pack ? (len + 3u) >> 2 : len; // output bytes
        if (nb == 0) continue;
        const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
        uint32_t *d0 = reinterpret_cast<uint32_t *>(dst - head); // word j holds output bytes [4j - head, 4j - head + 4)
        const uint32_t W = (head + nb + 3u) >> 2;
        auto out_byte = [&](uint32_t b) -> uint8_t { // one output byte the slow way (edge words)
            if (!A.pack) return src[b];
            unsigned v = 0;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++)
                if (b * 4 + t < len) v |= (unsigned)(src[b * 4 + t] & 3u) << (2 * t);
            return (uint8_t)v;
        };
        // edge words: lanes 0-3 the first word's bytes, lanes 4-7 the last word's
        {
            const uint32_t b = sub < 4 ? sub : 4u * (W - 1u) - head + (sub - 4u);
            const bool mine = sub < 4 ? (b + head < 4u) : (W > 1u && b >= 4u * (W - 1u) - head);
            if (mine && b < nb) dst[b] = out_byte(b);
        }
        if (!A.pack) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(src) - head; // source address of word 0's first byte
            const uint32_t sh = (uint32_t)(a & 3u) * 8u;
            const uint32_t *aw = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
            for (uint32_t j0 = 1; j0 + 1 < W; j0 += 16) { // two words per lane in flight
                const uint32_t j1 = j0 + sub, j2 = j1 + 8u;
                uint32_t lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
                if (j1 + 1 < W) {
                    lo1 = aw[j1];
                    if (sh) hi1 = aw[j1 + 1];
                }
                if (j2 + 1 < W) {
                    lo2 = aw[j2];
                    if (sh) hi2 = aw[j2 + 1];
                }
                if (j1 + 1 < W) d0[j1] = sh ? (lo1 >> sh) | (hi1 << (32u - sh)) : lo1;
                if (j2 + 1 < W) d0[j2] = sh ? (lo2 >> sh) | (hi2 << (32u - sh)) : lo2;
            }
        } else {
            auto pack4 = [](uint32_t w) -> uint32_t { // four ops, one per byte -> 8 bits
                w &= 0x03030303u;
                w |= w >> 6;
                w |= w >> 12;
                return w & 0xffu;
            };
            const uintptr_t a = reinterpret_cast<uintptr_t>(src) - 4u * head; // source address of word 0's first op
            const uint32_t ws = (uint32_t)(a >> 2) & 3u, sh = (uint32_t)(a & 3u) * 8u;
            const uint4 *aq = reinterpret_cast<const uint4 *>(a & ~(uintptr_t)15);
            for (uint32_t j = 1 + sub; j + 1 < W; j += 8) {
                const uint4 c0 = aq[j];
                uint4 c1 = make_uint4(0, 0, 0, 0);
                if (ws | sh) c1 = aq[j + 1];
                // y[0..4] = the five words from word offset ws on
                uint32_t y0 = c0.x, y1 = c0.y, y2 = c0.z, y3 = c0.w, y4 = c1.x;
                if (ws == 1) y0 = c0.y, y1 = c0.z, y2 = c0.w, y3 = c1.x, y4 = c1.y;
                if (ws == 2) y0 = c0.z, y1 = c0.w, y2 = c1.x, y3 = c1.y, y4 = c1.z;
                if (ws == 3) y0 = c0.w, y1 = c1.x, y2 = c1.y, y3 = c1.z, y4 = c1.w;
                if (sh) {
                    y0 = (y0 >> sh) | (y1 << (32u - sh));
                    y1 = (y1 >> sh) | (y2 << (32u - sh));
                    y2 = (y2 >> sh) | (y3 << (32u - sh));
                    y3 = (y3 >> sh) | (y4 << (32u - sh));
                }
                d0[j] = pack4(y0) | (pack4(y1) << 8) | (pack4(y2) << 16) | (pack4(y3) << 24);
            }
        }
    }
}

// ---- synthetic inputs: i.i.d. uniform DNA from a counter-based generator ----------------
// base(pos) of sequence w of pair p = "ACGT"[(splitmix64(key(p,w) + pos/32) >> (2*(pos%32))) & 3]
struct GenArgs {
    uint64_t seed, first_pair, n_pairs;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    uint8_t *bases;
};

__global__ void __launch_bounds__(256) generate_kernel(GenArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t p = gw; p < A.n_pairs; p += nw) {
        for (int w = 0; w < 2; w++) {
            const uint32_t len = w ? A.len2[p] : A.len1[p];
            uint8_t *dst = A.bases + (w ? A.off2[p] : A.off1[p]);
            const uint64_t key = synth_key(A.seed, A.first_pair + p, w);
            for (uint32_t pos = lane; pos < len; pos += 32) {
                const uint64_t word = splitmix64(key + (uint64_t)(pos >> 5));
                dst[pos] = (uint8_t)("ACGT"[(word >> (2 * (pos & 31))) & 3]);
            }
        }
    }
}

// ---- integer-pipe micro-benchmark (SURVEY.md 8d "Peak") -------------------------------------------------
// ILP independent dependency chains of one instruction class per thread; reports via clock64.
#ifndef SEQA_EMU
template <int WHICH>
__global__ void __launch_bounds__(256) int_peak_kernel(unsigned *out, int iters, unsigned seed, long long *cycles)
{
    constexpr int ILP = 8;
    unsigned x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        x[k] = seed * (k + 1) + threadIdx.x;
        y[k] = seed ^ (0x9E3779B9u * (k + 3));
    }
    const unsigned c = seed | 0x00010001u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < ILP; k++) {
                if (WHICH == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c));
                else if (WHICH == 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));
                else if (WHICH == 2) x[k] = (unsigned)__viaddmax_s32((int)x[k], (int)c, (int)y[k]);
                else if (WHICH == 3) x[k] = __vimax3_s16x2(x[k], y[k], c);
                else if (WHICH == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c), "r"(y[k]));
                else if (WHICH == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(c), "r"(y[k]));
                else if (WHICH == 6) x[k] = seqa_prmt(x[k], y[k], 0x6240);
                else if (WHICH == 7) x[k] = __vadd2(x[k], c);
                else if (WHICH == 8) x[k] = __viaddmax_s16x2_relu(x[k], c, y[k]);
                else { // 9: the packed SW cell mix: PRMT + VIADD.16x2 + VIADDMNMX.RELU + VIADDMNMX + VIMNMX
                    const unsigned sim = seqa_prmt(y[k], c, x[k]);
                    const unsigned lg = __vadd2(x[k], c);
                    const unsigned t = __viaddmax_s16x2_relu(y[k], sim, lg);
                    x[k] = __viaddmax_s16x2(x[k], c, t);
                    y[k] = __vmaxs2(y[k], x[k]);
                }
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc ^= x[k] ^ y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
#endif
