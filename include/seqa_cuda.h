/*
 * seqa_cuda.h -- C ABI of the B200 (sm_100a) batched pairwise-alignment library, libseqa_cuda.so.
 *
 * The reference (przemektmalon/SeqALib) is a header-only C++14 template library with NO foreign
 * interface; its plugin boundary is the virtual
 *     AlignedSequence<Ty,Blank> SequenceAligner::getAlignment(ContainerType&, ContainerType&)
 * (reference include/SequenceAlignment.h:133-153).  This header is the boundary a maintainer binds
 * underneath that virtual: every entry point below replaces the body of one reference function chain,
 * cited per declaration.  include/SequenceAlignment.h of THIS repository is the host-side mirror of the
 * reference's template API that calls these entry points; INTEGRATION.md shows the stub a maintainer
 * of the reference itself would add.
 *
 * Conventions
 *   * plain pointers and sizes only; the caller owns every buffer (pinned host memory recommended);
 *   * every function returns 0 (SEQA_OK) or a negative seqa_status, never throws, never aborts;
 *     seqa_cuda_last_error() returns a thread-local human-readable message for the last failure;
 *   * there is NO CPU fallback: with no usable CUDA device every compute call fails with
 *     SEQA_ERR_NO_DEVICE;
 *   * an alignment is reported as ops in FORWARD order (0 = diagonal (a_i,b_j), 1 = up (a_i,Blank),
 *     2 = left (Blank,b_j)), one byte per op, covering rows [start_i,end_i) x columns [start_j,end_j)
 *     of the DP matrix.  Global algorithms: start=(0,0), end=(len1,len2).  Local algorithms
 *     (SmithWaterman, LocalGotoh) report only the local part; the reference's forceGlobal framing
 *     (include/SequenceAlignment.h:156-189) is a pure function of the four indices and is applied by
 *     the host mirror.
 */
#ifndef SEQA_CUDA_H
#define SEQA_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEQA_ABI_VERSION 2

typedef enum seqa_algo {
    SEQA_NW = 0,          /* NeedlemanWunschSA   reference include/SANeedlemanWunsch.h:40-231 */
    SEQA_SW = 1,          /* SmithWatermanSA     reference include/SASmithWaterman.h:47-339   */
    SEQA_GLOBAL_GOTOH = 2,/* GlobalGotohSA       reference include/SAGlobalGotoh.h:53-422     */
    SEQA_LOCAL_GOTOH = 3, /* LocalGotohSA        reference include/SALocalGotoh.h:56-490      */
    SEQA_HIRSCHBERG = 4,  /* HirschbergSA        reference include/SAHirschberg.h:11-184      */
    SEQA_MYERS_MILLER = 5 /* MyersMillerSA       reference include/SAMyersMiller.h:43-420     */
} seqa_algo;

typedef enum seqa_status {
    SEQA_OK = 0,
    SEQA_ERR_INVALID = -1,     /* null pointer, bad enum, inconsistent sizes */
    SEQA_ERR_UNSUPPORTED = -2, /* scoring / shape outside the supported domain (see seqa_params) */
    SEQA_ERR_NO_DEVICE = -3,   /* no CUDA device / driver: there is no CPU fallback */
    SEQA_ERR_CUDA = -4,        /* a CUDA call failed; message in seqa_cuda_last_error() */
    SEQA_ERR_CAPACITY = -5,    /* an output buffer is too small (ops_capacity) */
    SEQA_ERR_NOMEM = -6        /* device or host allocation failed */
} seqa_status;

enum {
    SEQA_OP_DIAG = 0,
    SEQA_OP_UP = 1,
    SEQA_OP_LEFT = 2
};

/* seqa_params.flags */
#define SEQA_FLAG_SCORE_ONLY 0x1u  /* skip traceback/ops (scores and end positions only) */
#define SEQA_FLAG_FORCE_GENERIC 0x2u /* use the generic int32 kernels even where a packed fast path applies */
#define SEQA_FLAG_TRACE8 0x4u /* packed path: keep 8 trace bits per cell even where 4 or 2 suffice (testing) */
#define SEQA_FLAG_TRACE4 0x40u /* packed linear path: at least 4 trace bits per cell even where 2 suffice (testing) */
#define SEQA_FLAG_OPS_2BIT 0x10u /* ops leave the device packed 4 per byte: op k of pair p sits in bits 2*(k%4) of byte
                                   ops[ops_off[p] + k/4]; ops_off is in BYTES (every pair starts on a byte boundary), ops_len
                                   in OPS; ops_capacity >= sum(len1+len2)/4 + n_pairs suffices.  A quarter of the PCIe bytes
                                   of the default one-byte-per-op form ("2 bits each or one byte each (flag)") */
#define SEQA_FLAG_BASES_2BIT 0x20u /* INPUT wire format: seqa_batch_in.bases holds 2-bit symbols, 4 per byte (symbol k of a
                                     sequence in bits 2*(k%4) of its byte k/4), code A = 0, C = 1, T = 2, G = 3 (= (letter >> 1) & 3);
                                     every sequence starts on a byte boundary; off1 / off2 are BYTE offsets into the packed
                                     buffer, len1 / len2 count symbols, bases_len counts bytes.  A quarter of the bytes cross
                                     PCIe; the device expands them once.  Only the letters ACGT have a code: a caller whose
                                     sequences hold anything else sends 8-bit symbols (flag clear).  north_star: "sequences
                                     packed 2-bit/8-bit"; replaces the same (Seq1, Seq2) arguments as the 8-bit form */
#define SEQA_FLAG_LS_R1 0x8u /* linear-space path: 32-row blocks everywhere (testing: deep row-block pipelines on short pairs) */

/*
 * Mirror of the reference ScoringSystem (include/SequenceAlignment.h:82-131) plus the algorithm and
 * the device range.  Supported domain: gap < 0 (linear algorithms),
 * gap_open <= 0 and gap_extend < 0 (affine algorithms), match > 0, mismatch < 0 when
 * allow_mismatch != 0.  allow_mismatch == 0 reproduces the reference's "mismatch = INT_MIN constant"
 * behaviour (include/SANeedlemanWunsch.h:55-57,138); `mismatch` is then ignored.
 * LocalGotoh shapes (314,288), (60,57), (61,58) hit undefined behaviour in the reference
 * (include/SALocalGotoh.h:484-488: the result is replaced by NeedlemanWunsch with an uninitialised Gap) and are
 * rejected PER PAIR: the call succeeds, the pair comes back with ops_len = SEQA_PAIR_UNSUPPORTED, score = INT32_MIN,
 * start = end = (0,0) and no ops; every other pair of the batch is aligned normally.
 */
#define SEQA_PAIR_UNSUPPORTED 0xffffffffu /* seqa_batch_out.ops_len of a pair the GPU path rejects */
typedef struct seqa_params {
    int32_t algo; /* seqa_algo */
    int32_t gap;
    int32_t gap_open;
    int32_t gap_extend;
    int32_t match;
    int32_t mismatch;
    int32_t allow_mismatch;
    int32_t device_first; /* first CUDA ordinal to use */
    int32_t device_count; /* number of consecutive devices; 0 = all visible from device_first */
    uint32_t flags;
} seqa_params;

/*
 * A batch of independent pairs.  Sequence w of pair p is bases[offw[p] .. offw[p]+lenw[p]), 8-bit
 * symbols compared with == (the reference's nullptr-functor path, e.g. include/SANeedlemanWunsch.h:113).
 * Replaces the (ContainerType& Seq1, ContainerType& Seq2) arguments of getAlignment.
 */
typedef struct seqa_batch_in {
    const char *bases;
    const uint64_t *off1;
    const uint64_t *off2;
    const uint32_t *len1;
    const uint32_t *len2;
    uint64_t n_pairs;
    uint64_t bases_len; /* total bytes addressable through `bases` */
    /* Table-driven symbol equality, optional: NULL = symbols match iff they are the same byte.
     * Else 256 entries: two symbols a, b match iff sym_class[a] == sym_class[b] and that class is not SEQA_CLASS_NEVER --
     * the matching functors that ARE an equivalence on bytes (case-insensitive letters, purine / pyrimidine, amino-acid
     * groups; 'N' that matches nothing: SEQA_CLASS_NEVER) without a functor call per cell (the reference evaluates
     * match(a, b) per cell, include/SequenceAlignment.h:147 and e.g. include/SANeedlemanWunsch.h:22-38).  Class ids are
     * 0..253.  The device replaces every symbol by a representative of its class once; the classes of 'A', 'C', 'G', 'T'
     * keep the packed s16x2 kernels, pairs holding any other class run on the 8-bit kernels (pair by pair). */
    const uint8_t *sym_class;
} seqa_batch_in;
#define SEQA_CLASS_NEVER 255u

/*
 * Results, one entry per pair (arrays of n_pairs elements; any of the start/end/ops arrays may be NULL when
 * SEQA_FLAG_SCORE_ONLY is set).  ops for pair p are ops[ops_off[p] .. ops_off[p]+ops_len[p]).
 * ops_capacity >= sum(len1+len2) always suffices.  ops_used receives the bytes written.
 * With SEQA_FLAG_OPS_2BIT the same ops are packed 4 per byte (see the flag).
 * Replaces the AlignedSequence<Ty,Blank> return value (include/SequenceAlignment.h:13-80).
 * score: NW/GlobalGotoh H[M][N]; SW/LocalGotoh MaxScore (0 for an empty input); Hirschberg/MyersMiller
 * (which expose no score in the reference) the score of the returned alignment under the algorithm's
 * own gap model (a run of k gaps costs gap_open + k*gap_extend for MyersMiller).
 */
typedef struct seqa_batch_out {
    int32_t *score;
    uint32_t *start_i;
    uint32_t *start_j;
    uint32_t *end_i;
    uint32_t *end_j;
    uint8_t *ops;
    uint64_t *ops_off;
    uint32_t *ops_len;
    uint64_t ops_capacity;
    uint64_t ops_used;
} seqa_batch_out;

/* ---- one-shot entry: host buffers in, host buffers out ------------------------------------------
 * Replaces cacheAllMatches + computeScoreMatrix + buildResult + clearAll of the chosen aligner
 * (e.g. reference include/SANeedlemanWunsch.h:256-264) for every pair of the batch.  Pairs are split
 * statically over the device range (balanced by sum len1*len2), one host thread per device, no
 * inter-device communication.  Re-entrant for disjoint device sets. */
int seqa_cuda_align_batch(const seqa_params *params, const seqa_batch_in *in, seqa_batch_out *out);

/* The same call for a caller that PACKS its batch while the GPU works: `len1` / `len2` / `off1` / `off2` must be complete
 * when the call starts (they plan the waves), but the symbols in `bases` may be produced lazily -- right before the library
 * reads the symbols of pairs [first_pair, first_pair + n_pairs) it calls fill(user, first_pair, n_pairs) from one of its own
 * threads (every pair exactly once, increasing order per device, concurrently for different devices).  The caller's packing of
 * wave k+1 then overlaps the upload, kernels and download of wave k inside ONE call.  fill returns 0, or a non-zero value to
 * abort the call (SEQA_ERR_INVALID; e.g. a symbol the chosen wire format cannot hold).  include/SequenceAlignment.h packs its
 * std::string pairs this way. */
typedef int (*seqa_fill_fn)(void *user, uint64_t first_pair, uint64_t n_pairs);
int seqa_cuda_align_batch_lazy(const seqa_params *params, const seqa_batch_in *in, seqa_batch_out *out, seqa_fill_fn fill, void *user);

/* Frees the per-device contexts (device buffers) that seqa_cuda_align_batch creates lazily and keeps between
 * calls -- the only hidden state of the library. */
void seqa_cuda_trim(void);

/* Page-locked host memory for the caller's batch buffers (cudaHostAlloc, portable across devices): transfers from / to
 * such buffers run at PCIe speed and overlap with the kernels; pageable buffers work too but are staged by the
 * driver at a fraction of that.  NULL on failure (message in seqa_cuda_last_error()).  include/SequenceAlignment.h
 * keeps its packing buffers in memory from here. */
void *seqa_cuda_host_alloc(uint64_t bytes);
void seqa_cuda_host_free(void *ptr);

/* How the calling thread's last seqa_cuda_align_batch split its batch: sum(len1*len2 + 1) per device of the range
 * (the static split of SURVEY.md 8e, balanced by cells).  Returns the number of devices used; fills at most
 * `capacity` entries. */
int seqa_cuda_last_split(uint64_t *cells_per_device, int32_t capacity);

const char *seqa_cuda_last_error(void);
int seqa_cuda_device_count(void); /* number of visible CUDA devices, 0 if none / no driver */
int seqa_cuda_abi_version(void);

/* ---- resident (device-side) interface ------------------------------------------------------------
 * The same path split into its stages so that a caller (bench.py, a GPU-resident producer/consumer)
 * can keep inputs and results in HBM: create -> upload | generate -> run (repeatable) -> download.
 * `stream` is a cudaStream_t the caller owns (the caller's own stream), or NULL for a private stream. */
typedef struct seqa_ctx seqa_ctx;

int seqa_ctx_create(seqa_ctx **ctx, int device, void *stream);
void seqa_ctx_destroy(seqa_ctx *ctx);
/* Copy a batch to the device (asynchronous on the ctx stream when the host buffers are pinned). */
int seqa_ctx_upload(seqa_ctx *ctx, const seqa_params *params, const seqa_batch_in *in);
/* Fill the device with synthetic pairs from the shared counter-based generator of SURVEY.md 8d:
 * pair p = first_pair + k.  len_mode 0: fixed (len1,len2); 1: independent U[50,1000] per sequence
 * (len1/len2 ignored).  Nothing crosses PCIe. */
int seqa_ctx_generate(seqa_ctx *ctx, const seqa_params *params, uint64_t seed, uint64_t first_pair,
                      uint64_t n_pairs, int32_t len_mode, uint32_t len1, uint32_t len2);
/* DP fill + traceback (+ recursion for the linear-space algorithms) for the resident batch; results stay
 * on the device.  Asynchronous on the ctx stream. */
int seqa_ctx_run(seqa_ctx *ctx);
/* Copy the results of the last run to host buffers (synchronises the stream). */
int seqa_ctx_download(seqa_ctx *ctx, seqa_batch_out *out);
/* Results of pairs [first, first + count) of the last run only (arrays of `count` elements; ops_off rebased so that
 * the slice's first op string starts at out->ops[0]; ops_capacity >= sum(len1+len2) of the slice suffices): a
 * consumer that samples or streams the results of a large resident batch.  Synchronises the stream. */
int seqa_ctx_download_range(seqa_ctx *ctx, uint64_t first, uint64_t count, seqa_batch_out *out);
int seqa_ctx_sync(seqa_ctx *ctx);
/* Device-resident results: fills `dev` with DEVICE pointers to the arrays of the last run
 * (same layout and meaning as seqa_ctx_download would produce; ops are dense, ops_off relative to dev->ops) for a
 * GPU consumer on the ctx stream.  Synchronises once (the packed path's bad-symbol check); the pointers stay valid
 * until the next upload / generate / run on this ctx.  dev->ops_capacity and dev->ops_used receive the ops bytes. */
int seqa_ctx_device_results(seqa_ctx *ctx, seqa_batch_out *dev);
/* Number of kernels launched by this ctx since creation / sum(len1*len2) of the resident batch. */
uint64_t seqa_ctx_launch_count(const seqa_ctx *ctx);
uint64_t seqa_ctx_cells(const seqa_ctx *ctx);
/* Device time in ms of the dominant (DP fill) kernels of the last seqa_ctx_run, measured with CUDA events on
 * the ctx stream; valid after a sync.  *n_launches receives how many launches the sum covers. */
int seqa_ctx_last_fill_ms(seqa_ctx *ctx, float *ms, int32_t *n_launches);
/* Name of the dominant kernel family the last run used ("linear_s16x2", "linear_i32", ...). */
const char *seqa_ctx_last_kernel(const seqa_ctx *ctx);
/* Copy back the (possibly generated) resident inputs: bases as 8-bit characters in batch layout.
 * Buffers sized by the caller: bases_len >= sum(len1+len2), arrays of n_pairs. */
int seqa_ctx_download_inputs(seqa_ctx *ctx, char *bases, uint64_t bases_len, uint64_t *off1, uint64_t *off2,
                             uint32_t *len1, uint32_t *len2);

/* Integer-pipe micro-benchmark (SURVEY.md 8d "Peak"): runs dependency-free streams of one SASS
 * instruction class and reports lane-ops per clock per SM.  `which`: 0 IADD3, 1 VIMNMX(s32),
 * 2 VIADDMNMX(s32), 3 VIMNMX3.S16x2, 4 IMAD, 5 LOP3, 6 PRMT, 7 VIADD.16x2, 8 VIADDMNMX.S16x2.RELU,
 * 9 mix (ALU+FMA interleaved). */
int seqa_cuda_int_peak(int device, int which, double *lane_ops_per_clk_per_sm, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* SEQA_CUDA_H */
