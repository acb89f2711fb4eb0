// TEST INFRASTRUCTURE ONLY -- a tiny SIMT emulator so that the *unmodified* kernel sources under
// seqalib_b200/csrc/ can be compiled with g++ and executed on the CPU of the build container (which has
// no GPU).  It exists to debug kernel LOGIC (index arithmetic, tie-breaks, packed-integer tricks,
// warp-shuffle protocols) before GPU minutes are spent.  It is built only into tests/emu/libseqa_emu.so,
// is loaded only by `-m "not gpu"` tests, and is never loaded, linked or imported by the product
// (seqalib_b200/ loads libseqa_cuda.so and fails loudly without it).  It is NOT a CPU fallback.
//
// Model: every CUDA thread of a block is a ucontext fiber; the fibers of one block run round-robin on one
// OS thread and switch only inside collectives (__shfl_*_sync, __syncwarp, __syncthreads, __ballot_sync,
// ...), which are barriers over the warp / block.  Blocks of a grid run one after another.
#pragma once
#ifndef SEQA_EMU
#error "cuda_emu.h is only for the SEQA_EMU test build"
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include <ucontext.h>
#include <algorithm>
#include <vector>
#include <functional>
#include <chrono>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static thread_local
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r = {a, b}; return r; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }
static inline int2 make_int2(int a, int b) { int2 r = {a, b}; return r; }

// ---- fibers -------------------------------------------------------------------------------------------
struct EmuFiber {
    ucontext_t ctx;
    char *stack;
    uint3 tid;
    bool done;
};
struct EmuBlock {
    std::vector<EmuFiber> fibers;
    ucontext_t sched;
    int cur;
    uint3 bid;
    dim3 bdim, gdim;
    std::function<void()> body;
    // barrier state: index 0 = block barrier, 1+w = warp w
    std::vector<int> arrive;
    std::vector<unsigned> gen;
    std::vector<uint64_t> xch; // exchange slots, one per thread
    char *dyn_smem;
};
extern thread_local EmuBlock *emu_blk;

#define threadIdx (emu_blk->fibers[emu_blk->cur].tid)
#define blockIdx (emu_blk->bid)
#define blockDim (emu_blk->bdim)
#define gridDim (emu_blk->gdim)
#define warpSize 32

void emu_yield();
void emu_barrier(int group, int group_size);
void emu_run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
extern long emu_launches;

static inline int emu_lane() { return (int)(threadIdx.x & 31u); }
static inline int emu_warp() { return (int)(threadIdx.x >> 5); }
static inline int emu_warp_size() // number of live threads of my warp
{
    int w = emu_warp();
    int n = (int)blockDim.x - w * 32;
    return n > 32 ? 32 : n;
}

template <class T> static inline T emu_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "exchange size");
    const int w = emu_warp(), l = emu_lane();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu_blk->xch[w * 32 + l] = raw;
    emu_barrier(1 + w, emu_warp_size());
    T r = v;
    if (src_lane >= 0 && src_lane < emu_warp_size()) {
        uint64_t q = emu_blk->xch[w * 32 + src_lane];
        memcpy(&r, &q, sizeof(T));
    }
    emu_barrier(1 + w, emu_warp_size());
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    int l = emu_lane();
    return emu_exchange(v, l >= (int)d ? l - (int)d : -1);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
    int l = emu_lane();
    return emu_exchange(v, l + (int)d < 32 ? l + (int)d : -1);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, emu_lane() ^ m); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    const int w = emu_warp(), l = emu_lane();
    emu_blk->xch[w * 32 + l] = pred ? 1u : 0u;
    emu_barrier(1 + w, emu_warp_size());
    unsigned r = 0;
    for (int k = 0; k < emu_warp_size(); k++)
        if (emu_blk->xch[w * 32 + k]) r |= 1u << k;
    emu_barrier(1 + w, emu_warp_size());
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p)
{
    unsigned full = emu_warp_size() == 32 ? 0xffffffffu : ((1u << emu_warp_size()) - 1u);
    return __ballot_sync(m, p) == full;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_barrier(1 + emu_warp(), emu_warp_size()); }
static inline void __syncthreads() { emu_barrier(0, (int)blockDim.x); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline long long clock64()
{
    return (long long)std::chrono::steady_clock::now().time_since_epoch().count();
}

template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int *p, int v)
{
    int o = *p;
    if (v > o) *p = v;
    return o;
}
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v)
{
    unsigned long long o = *p;
    if (v > o) *p = v;
    return o;
}
template <class T> static inline T __ldg(const T *p) { return *p; }

// ---- packed-integer intrinsics used by the kernels (bit-exact models of the sm_100a instructions) ------
static inline int16_t emu_lo(unsigned v) { return (int16_t)(v & 0xffffu); }
static inline int16_t emu_hi(unsigned v) { return (int16_t)(v >> 16); }
static inline unsigned emu_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | (((unsigned)hi & 0xffffu) << 16); }
static inline unsigned __vadd2(unsigned a, unsigned b) { return emu_pack(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b)); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return emu_pack(emu_lo(a) - emu_lo(b), emu_hi(a) - emu_hi(b)); }
static inline unsigned __vmaxs2(unsigned a, unsigned b)
{
    return emu_pack(std::max(emu_lo(a), emu_lo(b)), std::max(emu_hi(a), emu_hi(b)));
}
static inline unsigned __vmins2(unsigned a, unsigned b)
{
    return emu_pack(std::min(emu_lo(a), emu_lo(b)), std::min(emu_hi(a), emu_hi(b)));
}
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c)
{
    return __vmaxs2(__vmaxs2(__vadd2(a, b), c), 0u);
}
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vimax3_s16x2_relu(unsigned a, unsigned b, unsigned c)
{
    return __vmaxs2(__vmaxs2(__vmaxs2(a, b), c), 0u);
}
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) { return __vmaxs2(__vmaxs2(a, b), 0u); }
static inline int __viaddmax_s32(int a, int b, int c) { return std::max((int)((unsigned)a + (unsigned)b), c); }
static inline int __viaddmax_s32_relu(int a, int b, int c)
{
    return std::max(std::max((int)((unsigned)a + (unsigned)b), c), 0);
}
static inline int __vimax3_s32(int a, int b, int c) { return std::max(std::max(a, b), c); }
static inline int __vimax3_s32_relu(int a, int b, int c) { return std::max(std::max(std::max(a, b), c), 0); }
// PRMT, default mode: selector nibble k picks byte (nibble & 7) of {b:a}; bit 3 replicates its sign bit.
static inline unsigned emu_prmt(unsigned a, unsigned b, unsigned s)
{
    const uint64_t src = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        const unsigned nib = (s >> (4 * k)) & 0xfu;
        unsigned byte = (unsigned)(src >> (8 * (nib & 7u))) & 0xffu;
        if (nib & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
        r |= byte << (8 * k);
    }
    return r;
}

// the CUDA intrinsic masks the selector (verified in the PTX nvcc 12.9 emits: and.b32 s, 0x7777)
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) { return emu_prmt(a, b, s & 0x7777u); }

// ---- runtime API subset ----------------------------------------------------------------------------------
typedef int cudaError_t;
typedef struct EmuStream *cudaStream_t;
typedef struct EmuEvent { std::chrono::steady_clock::time_point t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[64];
    int multiProcessorCount, clockRate, major, minor;
    size_t totalGlobalMem, sharedMemPerBlockOptin;
};
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 2; return cudaSuccess; } // two fake devices: exercises the shard path
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "seqa-emu");
    p->multiProcessorCount = 2;
    p->clockRate = 1000000;
    p->major = 10;
    p->totalGlobalMem = (size_t)2 << 30;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)1 << 30; *t = (size_t)2 << 30; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
enum { cudaHostAllocPortable = 1 };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }

extern char *emu_dyn_smem_ptr();
#define SEQA_DYN_SMEM(type, name) type *name = (type *)emu_dyn_smem_ptr()

// kernel launch: SEQA_LAUNCH((kernel<...>), grid, block, smem_bytes, stream, args...)
#define SEQA_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
    do {                                                                                     \
        (void)(stream);                                                                      \
        emu_run_grid(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); }); \
    } while (0)
