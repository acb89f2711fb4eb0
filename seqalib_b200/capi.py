"""ctypes binding of the C ABI in include/seqa_cuda.h (libseqa_cuda.so).

This is plumbing for tests and bench.py; the product host side is the C++ header include/SequenceAlignment.h.
The library is looked up in-tree (seqalib_b200/libseqa_cuda.so, built by `make`) and loading fails loudly when it
is missing: there is no Python or CPU fallback for the alignment path.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.path.join(HERE, "libseqa_cuda.so")

ALGOS = {"nw": 0, "sw": 1, "ggotoh": 2, "lgotoh": 3, "hirschberg": 4, "myersmiller": 5}
FLAG_SCORE_ONLY = 1
FLAG_FORCE_GENERIC = 2
FLAG_TRACE8 = 4
FLAG_TRACE4 = 64  # packed linear path: no 2-bit trace (testing)
FLAG_LS_R1 = 8
FLAG_OPS_2BIT = 16  # ops packed 4 per byte on the wire (Results.pair_ops unpacks)
FLAG_BASES_2BIT = 32  # INPUT symbols packed 4 per byte (A0 C1 T2 G3), see pack_bases_2bit
PAIR_UNSUPPORTED = 0xFFFFFFFF  # Results.ops_len of a pair the GPU path rejects (SEQA_PAIR_UNSUPPORTED)
OK = 0
ERR_NAMES = {0: "SEQA_OK", -1: "SEQA_ERR_INVALID", -2: "SEQA_ERR_UNSUPPORTED", -3: "SEQA_ERR_NO_DEVICE",
             -4: "SEQA_ERR_CUDA", -5: "SEQA_ERR_CAPACITY", -6: "SEQA_ERR_NOMEM"}

# every symbol include/seqa_cuda.h declares
EXPORTS = ["seqa_cuda_align_batch", "seqa_cuda_last_error", "seqa_cuda_device_count", "seqa_cuda_abi_version",
           "seqa_ctx_create", "seqa_ctx_destroy", "seqa_ctx_upload", "seqa_ctx_generate", "seqa_ctx_run",
           "seqa_ctx_download", "seqa_ctx_device_results", "seqa_ctx_sync", "seqa_ctx_launch_count", "seqa_ctx_cells", "seqa_ctx_last_fill_ms",
           "seqa_ctx_last_kernel", "seqa_ctx_download_inputs", "seqa_cuda_int_peak", "seqa_cuda_trim",
           "seqa_cuda_host_alloc", "seqa_cuda_host_free", "seqa_ctx_download_range", "seqa_cuda_last_split", "seqa_cuda_align_batch_lazy"]


def pack_bases_2bit(bases, off1, off2, len1, len2, out=None):
    """8-bit batch arrays -> the SEQA_FLAG_BASES_2BIT wire format: (packed u8, off1 u64, off2 u64) with 4 symbols per byte
    (code (letter >> 1) & 3: A0 C1 T2 G3), every sequence on a byte boundary, seq1 then seq2 per pair, pairs back to back.
    Raises ValueError when a symbol outside ACGT is present (such a batch must be sent as 8-bit symbols).
    Vectorised for uniform batches (bench.py's 1 M pairs); ragged batches take a per-symbol scatter (tests)."""
    n = len(len1)
    l1 = np.asarray(len1, dtype=np.int64)
    l2 = np.asarray(len2, dtype=np.int64)
    b1, b2 = (l1 + 3) // 4, (l2 + 3) // 4
    per = b1 + b2
    p1 = np.zeros(n, dtype=np.uint64)
    if n > 1:
        p1[1:] = np.cumsum(per)[:-1].astype(np.uint64)
    p2 = p1 + b1.astype(np.uint64)
    total = int(per.sum())
    packed = out if out is not None else np.zeros(max(total, 1), dtype=np.uint8)
    if n == 0:
        return packed, p1, p2
    letters = np.frombuffer(b"ACTG", dtype=np.uint8)
    uniform = bool((l1 == l1[0]).all() and (l2 == l2[0]).all())
    dense = uniform and bool((np.asarray(off1, dtype=np.int64) == np.arange(n, dtype=np.int64) * (l1[0] + l2[0])).all()) and \
        bool((np.asarray(off2, dtype=np.int64) == np.asarray(off1, dtype=np.int64) + l1[0]).all())
    if dense:
        L1, L2 = int(l1[0]), int(l2[0])
        view = np.asarray(bases)[:n * (L1 + L2)].reshape(n, L1 + L2)
        dst = packed[:total].reshape(n, int(per[0]))
        sh = np.array([0, 2, 4, 6], dtype=np.uint8)
        for (col, L, dcol, nb) in ((0, L1, 0, int(b1[0])), (L1, L2, int(b1[0]), int(b2[0]))):
            if L == 0:
                continue
            seg = view[:, col:col + L]
            code = (seg >> 1) & 3
            if not np.array_equal(letters[code], seg):
                raise ValueError("symbol outside ACGT: send this batch as 8-bit symbols")
            pad = np.zeros((n, nb * 4), dtype=np.uint8)
            pad[:, :L] = code
            dst[:, dcol:dcol + nb] = np.bitwise_or.reduce(pad.reshape(n, nb, 4) << sh[None, None, :], axis=2)
        return packed, p1, p2
    packed[:total] = 0
    for (off, ln, po) in ((off1, l1, p1), (off2, l2, p2)):
        tot = int(ln.sum())
        if tot == 0:
            continue
        owner = np.repeat(np.arange(n, dtype=np.int64), ln)
        within = np.arange(tot, dtype=np.int64) - (np.cumsum(ln) - ln)[owner]
        sym = np.asarray(bases)[np.asarray(off, dtype=np.int64)[owner] + within]
        code = (sym >> 1) & 3
        if not np.array_equal(letters[code], sym):
            raise ValueError("symbol outside ACGT: send this batch as 8-bit symbols")
        np.bitwise_or.at(packed, po.astype(np.int64)[owner] + within // 4, (code << (2 * (within % 4)).astype(np.uint8)).astype(np.uint8))
    return packed, p1, p2


class SeqaError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("algo", C.c_int32), ("gap", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("match", C.c_int32), ("mismatch", C.c_int32), ("allow_mismatch", C.c_int32),
                ("device_first", C.c_int32), ("device_count", C.c_int32), ("flags", C.c_uint32)]


class BatchIn(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("off1", C.c_void_p), ("off2", C.c_void_p), ("len1", C.c_void_p),
                ("len2", C.c_void_p), ("n_pairs", C.c_uint64), ("bases_len", C.c_uint64), ("sym_class", C.c_void_p)]


class BatchOut(C.Structure):
    _fields_ = [("score", C.c_void_p), ("start_i", C.c_void_p), ("start_j", C.c_void_p), ("end_i", C.c_void_p),
                ("end_j", C.c_void_p), ("ops", C.c_void_p), ("ops_off", C.c_void_p), ("ops_len", C.c_void_p),
                ("ops_capacity", C.c_uint64), ("ops_used", C.c_uint64)]


def make_params(algo, gap=0, gap_open=0, gap_extend=0, match=1, mismatch=-1, allow=True, device_first=0,
                device_count=1, flags=0):
    a = ALGOS[algo] if isinstance(algo, str) else int(algo)
    return Params(a, gap, gap_open, gap_extend, match, mismatch if allow else 0, 1 if allow else 0, device_first,
                  device_count, flags)


class Results(object):
    """Host copies of a seqa_batch_out."""

    def __init__(self, n, ops_capacity, pinned=None):
        alloc = pinned if pinned is not None else (lambda shape, dt: np.zeros(shape, dtype=dt))
        self.n = n
        self.score = alloc(max(n, 1), np.int32)
        self.start_i = alloc(max(n, 1), np.uint32)
        self.start_j = alloc(max(n, 1), np.uint32)
        self.end_i = alloc(max(n, 1), np.uint32)
        self.end_j = alloc(max(n, 1), np.uint32)
        self.ops_off = alloc(max(n, 1), np.uint64)
        self.ops_len = alloc(max(n, 1), np.uint32)
        self.ops = alloc(max(int(ops_capacity), 1), np.uint8)
        self.c = BatchOut(self.score.ctypes.data, self.start_i.ctypes.data, self.start_j.ctypes.data,
                          self.end_i.ctypes.data, self.end_j.ctypes.data, self.ops.ctypes.data,
                          self.ops_off.ctypes.data, self.ops_len.ctypes.data, int(ops_capacity), 0)

    packed2 = False  # set by the call that filled this object when SEQA_FLAG_OPS_2BIT was on

    def pair_ops(self, p):
        o, l = int(self.ops_off[p]), int(self.ops_len[p])
        if self.packed2:
            b = self.ops[o:o + (l + 3) // 4]
            return ((b[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)[None, :]) & 3).reshape(-1)[:l].astype(np.uint8)
        return self.ops[o:o + l]


class Lib(object):
    def __init__(self, path=None):
        path = path or os.environ.get("SEQA_LIB") or DEFAULT_SO  # SEQA_LIB: an experimental build of the same library (A/B runs)
        if not os.path.exists(path):
            raise ImportError("%s is missing: run `make` (nvcc, sm_100a) at the repository root. "
                              "There is no CPU fallback for the alignment path." % path)
        self.path = path
        L = self.L = C.CDLL(path)
        L.seqa_cuda_last_error.restype = C.c_char_p
        L.seqa_ctx_last_kernel.restype = C.c_char_p
        L.seqa_ctx_last_kernel.argtypes = [C.c_void_p]
        L.seqa_cuda_align_batch.argtypes = [C.POINTER(Params), C.POINTER(BatchIn), C.POINTER(BatchOut)]
        L.seqa_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        L.seqa_ctx_destroy.argtypes = [C.c_void_p]
        L.seqa_ctx_destroy.restype = None
        L.seqa_ctx_upload.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(BatchIn)]
        L.seqa_ctx_generate.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32,
                                        C.c_uint32, C.c_uint32]
        L.seqa_ctx_run.argtypes = [C.c_void_p]
        L.seqa_ctx_sync.argtypes = [C.c_void_p]
        L.seqa_ctx_download.argtypes = [C.c_void_p, C.POINTER(BatchOut)]
        L.seqa_ctx_device_results.argtypes = [C.c_void_p, C.POINTER(BatchOut)]
        L.seqa_ctx_download_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(BatchOut)]
        L.seqa_cuda_last_split.argtypes = [C.c_void_p, C.c_int32]
        L.seqa_ctx_launch_count.argtypes = [C.c_void_p]
        L.seqa_ctx_launch_count.restype = C.c_uint64
        L.seqa_ctx_cells.argtypes = [C.c_void_p]
        L.seqa_ctx_cells.restype = C.c_uint64
        L.seqa_ctx_last_fill_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.seqa_ctx_download_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p]
        L.seqa_cuda_trim.restype = None
        L.seqa_cuda_host_alloc.restype = C.c_void_p
        L.seqa_cuda_host_alloc.argtypes = [C.c_uint64]
        L.seqa_cuda_host_free.restype = None
        L.seqa_cuda_host_free.argtypes = [C.c_void_p]
        L.seqa_cuda_int_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]

    def check(self, rc):
        if rc != OK:
            raise SeqaError(rc, (self.L.seqa_cuda_last_error() or b"").decode("utf8", "replace"))

    def device_count(self):
        return self.L.seqa_cuda_device_count()

    def last_split(self):
        """cells per device of this thread's last align_batch call (seqa_cuda_last_split)."""
        buf = np.zeros(64, dtype=np.uint64)
        nd = self.L.seqa_cuda_last_split(buf.ctypes.data, 64)
        return buf[:nd].copy()

    @staticmethod
    def batch_in(bases, off1, off2, len1, len2, sym_class=None):
        """sym_class: optional uint8[256] class table (seqa_batch_in.sym_class); the caller keeps it alive during the call"""
        return BatchIn(bases.ctypes.data, off1.ctypes.data, off2.ctypes.data, len1.ctypes.data, len2.ctypes.data,
                       len(len1), len(bases), sym_class.ctypes.data if sym_class is not None else None)

    def align_batch(self, params, bases, off1, off2, len1, len2, results=None, sym_class=None):
        """seqa_cuda_align_batch: host arrays in (numpy, C-contiguous), Results out."""
        n = len(len1)
        if results is None:
            cap = int(len1.astype(np.uint64).sum() + len2.astype(np.uint64).sum())
            results = Results(n, cap)
        bi = self.batch_in(bases, off1, off2, len1, len2, sym_class)
        self.check(self.L.seqa_cuda_align_batch(C.byref(params), C.byref(bi), C.byref(results.c)))
        results.packed2 = bool(params.flags & FLAG_OPS_2BIT)
        return results


class Ctx(object):
    """Resident interface (seqa_ctx_*)."""

    def __init__(self, lib, device=0, stream=None):
        self.lib = lib
        self.h = C.c_void_p()
        lib.check(lib.L.seqa_ctx_create(C.byref(self.h), device, stream))
        self.n = 0
        self.slots = 0

    def close(self):
        if self.h:
            self.lib.L.seqa_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, params, bases, off1, off2, len1, len2, sym_class=None):
        bi = Lib.batch_in(bases, off1, off2, len1, len2, sym_class)
        self.lib.check(self.lib.L.seqa_ctx_upload(self.h, C.byref(params), C.byref(bi)))
        self.flags = int(params.flags)
        self.n = len(len1)
        self.slots = int(len1.astype(np.uint64).sum() + len2.astype(np.uint64).sum())

    def generate(self, params, seed, first_pair, n_pairs, len_mode=0, len1=0, len2=0):
        self.lib.check(self.lib.L.seqa_ctx_generate(self.h, C.byref(params), seed, first_pair, n_pairs, len_mode, len1, len2))
        self.flags = int(params.flags)
        self.n = n_pairs
        self.slots = None

    def run(self):
        self.lib.check(self.lib.L.seqa_ctx_run(self.h))

    def sync(self):
        self.lib.check(self.lib.L.seqa_ctx_sync(self.h))

    def download(self, results=None, ops_capacity=None):
        if results is None:
            results = Results(self.n, ops_capacity if ops_capacity is not None else self.slots)
        self.lib.check(self.lib.L.seqa_ctx_download(self.h, C.byref(results.c)))
        results.packed2 = bool(getattr(self, "flags", 0) & FLAG_OPS_2BIT)
        return results

    def download_range(self, first, count, ops_capacity):
        """Results of pairs [first, first+count) of the last run (seqa_ctx_download_range)."""
        results = Results(count, ops_capacity)
        self.lib.check(self.lib.L.seqa_ctx_download_range(self.h, first, count, C.byref(results.c)))
        results.packed2 = bool(getattr(self, "flags", 0) & FLAG_OPS_2BIT)
        return results

    def device_results(self):
        """Device pointers of the last run's results (a BatchOut whose pointer fields are device addresses)."""
        out = BatchOut()
        self.lib.check(self.lib.L.seqa_ctx_device_results(self.h, C.byref(out)))
        return out

    def download_inputs(self, total_bases):
        bases = np.zeros(max(total_bases, 1), dtype=np.uint8)
        off1 = np.zeros(max(self.n, 1), dtype=np.uint64)
        off2 = np.zeros(max(self.n, 1), dtype=np.uint64)
        len1 = np.zeros(max(self.n, 1), dtype=np.uint32)
        len2 = np.zeros(max(self.n, 1), dtype=np.uint32)
        self.lib.check(self.lib.L.seqa_ctx_download_inputs(self.h, bases.ctypes.data, total_bases, off1.ctypes.data,
                                                           off2.ctypes.data, len1.ctypes.data, len2.ctypes.data))
        return bases, off1[:self.n], off2[:self.n], len1[:self.n], len2[:self.n]

    def cells(self):
        return int(self.lib.L.seqa_ctx_cells(self.h))

    def launch_count(self):
        return int(self.lib.L.seqa_ctx_launch_count(self.h))

    def last_fill_ms(self):
        ms, nl = C.c_float(), C.c_int32()
        self.lib.check(self.lib.L.seqa_ctx_last_fill_ms(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def last_kernel(self):
        return (self.lib.L.seqa_ctx_last_kernel(self.h) or b"").decode()
