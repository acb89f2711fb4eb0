// Shared device/host definitions for libseqa_cuda.so (sm_100a only).
#pragma once
#include <stdint.h>
#include <limits.h>
#ifdef SEQA_EMU
// Test-only build of the same sources for the SIMT emulator (tests/emu/); never shipped.
#include "cuda_emu.h"
using std::max;
using std::min;
#else
#include <cuda_runtime.h>
// Kernel launch through one macro so the sources also build for the emulator: the kernel name is passed in
// parentheses, e.g. SEQA_LAUNCH((fill_i32_kernel<true, false, 4>), grid, block, smem, stream, args).
#define SEQA_LAUNCH(kern, grid, block, smem, stream, ...)        \
    do {                                                          \
        auto seqa_kfn_ = kern;                                    \
        seqa_kfn_<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    } while (0)
#define SEQA_DYN_SMEM(type, name)                                      \
    extern __shared__ __align__(16) unsigned char seqa_dyn_smem_raw_[]; \
    type *name = reinterpret_cast<type *>(seqa_dyn_smem_raw_)
#endif

// PRMT in its native default mode: selector nibble k picks byte (nibble & 7) of {b:a}; nibble bit 3 replicates
// that byte's sign bit instead.  (The CUDA intrinsic __byte_perm masks the selector with 0x7777 -- an extra LOP3
// and no sign replication -- so the kernels go to the PTX instruction directly.)
#ifdef SEQA_EMU
static inline unsigned seqa_prmt(unsigned a, unsigned b, unsigned s) { return emu_prmt(a, b, s); }
#else
__device__ __forceinline__ unsigned seqa_prmt(unsigned a, unsigned b, unsigned s)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
#endif

// ---- TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP) and the mbarrier its bytes arrive on ----
// One barrier per warp (arrival count 1): the issuing lane arms it with the byte count and issues the copy, every
// lane of the warp polls the phase parity.  Source and destination 16-byte aligned, size a multiple of 16.
#ifdef SEQA_EMU
static inline void seqa_mbar_init(uint64_t *bar) { *bar = 0; }
static inline void seqa_bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *) { memcpy(dst, src, bytes); }
static inline void seqa_mbar_wait(uint64_t *, unsigned) { __syncwarp(); } // the emulated copy is lane 0's memcpy
#else
__device__ __forceinline__ unsigned seqa_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void seqa_mbar_init(uint64_t *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(seqa_smem_addr(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void seqa_bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    const unsigned b = seqa_smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(seqa_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void seqa_mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned b = seqa_smem_addr(bar);
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(b), "r"(parity)
                     : "memory");
    } while (!done);
}
#endif

#define SEQA_WARP 32
#define SEQA_FULL 0xffffffffu
#define SEQA_GOTOH_NEG (-10000) /* the reference's literal "-infinity", include/SAGlobalGotoh.h:78-79 */

// Scoring constants as the kernels see them (reference include/SequenceAlignment.h:82-131).
struct DevScoring {
    int gap;        // linear gap penalty
    int go, ge;     // affine open / extend
    int match;
    int mismatch;   // only read when allow != 0
    int allow;      // AllowMismatch
};

// The diagonal candidate of every reference recurrence: H[i-1][j-1] + sim when mismatches are allowed,
// else the CONSTANT INT_MIN for a non-matching cell (include/SANeedlemanWunsch.h:117-118 vs :138).
__device__ __forceinline__ int diag_cand(const DevScoring &s, int hdiag, bool eq)
{
    if (s.allow) return hdiag + (eq ? s.match : s.mismatch);
    return eq ? hdiag + s.match : INT_MIN;
}

// Borders of one DP sweep, as affine functions of the index (index >= 1; H(0,0) is always 0):
//   H(i,0) = hcolA + i*hcolB      H(0,j) = hrowA + j*hrowB
//   Ix(0,j) = ixA + j*ixB (vertical-gap state of row 0)   Iy(i,0) = iyA + i*iyB (horizontal-gap state of column 0)
struct Borders {
    int hcolA, hcolB, hrowA, hrowB;
    int ixA, ixB, iyA, iyB;
};

__device__ __forceinline__ int border_hcol(const Borders &b, int i) { return i == 0 ? 0 : b.hcolA + i * b.hcolB; }
__device__ __forceinline__ int border_hrow(const Borders &b, int j) { return j == 0 ? 0 : b.hrowA + j * b.hrowB; }

// One sub-problem of the linear-space recursions (Hirschberg include/SAHirschberg.h:102-163,
// Myers-Miller include/SAMyersMiller.h:43-397): rows [i0,i0+m) x columns [j0,j0+n) of pair `pair`.
// q is the node's heap index inside its level (root 0, children 2q / 2q+1): rows of scratch for node q
// start at column offset j0 + q, which keeps the (n+1)-wide row buffers of one level disjoint.
struct LsNode {
    int pair;
    int i0, m;
    int j0, n;
    int q;
    int tb, te; // Myers-Miller boundary gap-open charges
};

// splitmix64 and the shared synthetic-input generator.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_key(uint64_t seed, uint64_t pair, int which)
{
    return splitmix64(seed ^ (2ull * pair + (uint64_t)which));
}
__host__ __device__ __forceinline__ uint32_t synth_len(uint64_t seed, uint64_t pair, int which)
{
    return 50u + (uint32_t)(splitmix64(synth_key(seed, pair, which) ^ 0xC0FFEEull) % 951ull);
}
