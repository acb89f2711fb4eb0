"""TEST INFRASTRUCTURE ONLY.  ctypes bindings for the two checkers under oracle/:

* ``libseqa_oracle.so`` -- the plain-C restatement (oracle/seqa_oracle.c), always available;
* ``_ref/libseqa_ref.so`` -- the unmodified reference compiled by path (oracle/ref_driver.cpp),
  available where it was prebuilt.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  Nothing in seqalib_b200/ imports it.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libseqa_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libseqa_ref.so")
REF_INCLUDE = "/root/reference/include"

ALGOS = {"nw": 0, "sw": 1, "ggotoh": 2, "lgotoh": 3, "hirschberg": 4, "myersmiller": 5}
ALGO_NAMES = {v: k for k, v in ALGOS.items()}
INT_MIN = -(2 ** 31)
OP_DIAG, OP_UP, OP_LEFT = 0, 1, 2


def build(force=False):
    """Compile the checkers (building the checker is not using it)."""
    src = os.path.join(HERE, "seqa_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libseqa_oracle.so"])
    if os.path.isdir(REF_INCLUDE):
        drv = os.path.join(HERE, "ref_driver.cpp")
        if force or not os.path.exists(REF_SO) or os.path.getmtime(REF_SO) < os.path.getmtime(drv):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Scoring(object):
    """Mirror of the reference ScoringSystem ctors (include/SequenceAlignment.h:92-118).

    ctor = 2: (gap, match) -> allow_mismatch False, mismatch INT_MIN
    ctor = 4: (gap, match, mismatch, allow)
    ctor = 5: (gap_open, gap_extend, match, mismatch, allow)
    """

    def __init__(self, ctor, gap=0, gap_open=0, gap_extend=0, match=0, mismatch=0, allow=True):
        self.ctor, self.gap, self.gap_open, self.gap_extend = ctor, gap, gap_open, gap_extend
        self.match, self.mismatch, self.allow = match, mismatch, bool(allow)
        if ctor == 2:
            self.mismatch, self.allow = INT_MIN, False

    @staticmethod
    def linear(gap, match, mismatch=None, allow=True):
        if mismatch is None:
            return Scoring(2, gap=gap, match=match)
        return Scoring(4, gap=gap, match=match, mismatch=mismatch, allow=allow)

    @staticmethod
    def affine(gap_open, gap_extend, match, mismatch, allow=True):
        return Scoring(5, gap_open=gap_open, gap_extend=gap_extend, match=match, mismatch=mismatch, allow=allow)

    def astuple(self):
        return (self.ctor, self.gap, self.gap_open, self.gap_extend, self.match, self.mismatch, int(self.allow))

    def __repr__(self):
        return "Scoring%r" % (self.astuple(),)


_ref = None
_orc = None


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_align.restype = C.c_int
        lib.ref_align.argtypes = [C.c_int] * 9 + [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                                   C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_bench.restype = C.c_double
        lib.ref_bench.argtypes = [C.c_int] * 8 + [C.c_void_p] * 5 + [C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        lib.ref_hardware_threads.restype = C.c_int
        _ref = lib
    return _ref


def _b(s):
    return s.encode("latin1") if isinstance(s, str) else bytes(s)


def ref_align(algo, sc, s1, s2, functor=False):
    """Run the unmodified reference.  Returns dict(row1,row2,flags,score,max_row,max_col)."""
    lib = ref_lib()
    a = ALGOS[algo] if isinstance(algo, str) else algo
    b1, b2 = _b(s1), _b(s2)
    cap = 2 * (len(b1) + len(b2)) + 8
    r1, r2, fl = C.create_string_buffer(cap), C.create_string_buffer(cap), C.create_string_buffer(cap)
    sc_, mr, mc = C.c_int(), C.c_int(), C.c_int()
    ctor, gap, go, ge, m, x, allow = sc.astuple()
    n = lib.ref_align(a, ctor, gap, go, ge, m, x, allow, int(functor), b1, len(b1), b2, len(b2),
                      r1, r2, fl, cap, C.byref(sc_), C.byref(mr), C.byref(mc))
    assert n <= cap
    return dict(row1=r1.raw[:n].decode("latin1"), row2=r2.raw[:n].decode("latin1"),
                flags="".join("|" if c else " " for c in fl.raw[:n]),
                score=None if sc_.value == INT_MIN else sc_.value, max_row=mr.value, max_col=mc.value)


def batch_arrays(pairs):
    """list of (seq1, seq2) -> (bases u8, off1 u64, off2 u64, len1 u32, len2 u32) in the C-ABI batch layout."""
    l1 = np.array([len(a) for a, _ in pairs], dtype=np.uint32)
    l2 = np.array([len(b) for _, b in pairs], dtype=np.uint32)
    blob = b"".join(_b(a) + _b(b) for a, b in pairs)
    tot = (l1.astype(np.uint64) + l2.astype(np.uint64))
    off1 = np.zeros(len(pairs), dtype=np.uint64)
    if len(pairs) > 1:
        off1[1:] = np.cumsum(tot)[:-1]
    off2 = off1 + l1.astype(np.uint64)
    bases = np.frombuffer(blob, dtype=np.uint8).copy() if blob else np.zeros(1, dtype=np.uint8)
    return bases, off1, off2, l1, l2


def ref_bench(algo, sc, bases, off1, off2, len1, len2, threads):
    """Threaded wall-clock run of the reference getAlignment over a batch -> (seconds, entries)."""
    lib = ref_lib()
    a = ALGOS[algo] if isinstance(algo, str) else algo
    ctor, gap, go, ge, m, x, allow = sc.astuple()
    ent = C.c_uint64()
    sec = lib.ref_bench(a, ctor, gap, go, ge, m, x, allow, bases.ctypes.data, off1.ctypes.data, off2.ctypes.data,
                        len1.ctypes.data, len2.ctypes.data, len(len1), threads, C.byref(ent))
    return sec, ent.value


def oracle_lib():
    global _orc
    if _orc is None:
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_align.restype = C.c_int
        lib.oracle_align.argtypes = [C.c_int] * 7 + [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                                      C.c_void_p, C.c_int, C.POINTER(C.c_int * 5)]
        lib.oracle_bench.restype = C.c_double
        lib.oracle_bench.argtypes = [C.c_int] * 7 + [C.c_void_p] * 5 + [C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        lib.oracle_align_batch.restype = C.c_int
        lib.oracle_align_batch.argtypes = [C.c_int] * 7 + [C.c_void_p] * 5 + [C.c_uint64, C.c_int] + [C.c_void_p] * 8
        _orc = lib
    return _orc


def oracle_align(algo, sc, s1, s2):
    """Run the C restatement.  Returns dict(score,start_i,start_j,end_i,end_j,ops[np.uint8 forward])."""
    lib = oracle_lib()
    a = ALGOS[algo] if isinstance(algo, str) else algo
    b1, b2 = _b(s1), _b(s2)
    cap = len(b1) + len(b2) + 8
    ops = np.zeros(cap, dtype=np.uint8)
    meta = (C.c_int * 5)()
    _, gap, go, ge, m, x, allow = sc.astuple()
    n = lib.oracle_align(a, gap, go, ge, m, x, allow, b1, len(b1), b2, len(b2), ops.ctypes.data, cap, C.byref(meta))
    if n < 0:
        raise RuntimeError("oracle_align failed: %d" % n)
    return dict(score=meta[0], start_i=meta[1], start_j=meta[2], end_i=meta[3], end_j=meta[4], ops=ops[:n].copy())


class OracleBatch:
    """Results of oracle_align_batch in the C-ABI's array layout (ops of pair p at slot_off[p], length ops_len[p])."""

    def pair_ops(self, p):
        o = int(self.slot_off[p])
        return self.ops[o:o + int(self.ops_len[p])]


def oracle_align_batch(algo, sc, bases, off1, off2, len1, len2, threads=0):
    """Threaded C-oracle run over a whole batch -> OracleBatch (for the >= 100 k-pair parity samples)."""
    lib = oracle_lib()
    a = ALGOS[algo] if isinstance(algo, str) else algo
    _, gap, go, ge, m, x, allow = sc.astuple()
    n = len(len1)
    threads = threads or len(os.sched_getaffinity(0))
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off1 = np.ascontiguousarray(off1, dtype=np.uint64)
    off2 = np.ascontiguousarray(off2, dtype=np.uint64)
    len1 = np.ascontiguousarray(len1, dtype=np.uint32)
    len2 = np.ascontiguousarray(len2, dtype=np.uint32)
    r = OracleBatch()
    slots = len1.astype(np.uint64) + len2
    r.slot_off = np.zeros(n, dtype=np.uint64)
    if n > 1:
        np.cumsum(slots[:-1], out=r.slot_off[1:])
    r.score = np.zeros(n, dtype=np.int32)
    r.start_i, r.start_j, r.end_i, r.end_j, r.ops_len = (np.zeros(n, dtype=np.uint32) for _ in range(5))
    r.ops = np.zeros(int(slots.sum()) + 8, dtype=np.uint8)
    rc = lib.oracle_align_batch(a, gap, go, ge, m, x, allow, bases.ctypes.data, off1.ctypes.data, off2.ctypes.data,
                                len1.ctypes.data, len2.ctypes.data, n, threads, r.score.ctypes.data, r.start_i.ctypes.data,
                                r.start_j.ctypes.data, r.end_i.ctypes.data, r.end_j.ctypes.data, r.ops_len.ctypes.data,
                                r.slot_off.ctypes.data, r.ops.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_align_batch failed: %d" % rc)
    return r


def oracle_bench(algo, sc, bases, off1, off2, len1, len2, threads):
    lib = oracle_lib()
    a = ALGOS[algo] if isinstance(algo, str) else algo
    _, gap, go, ge, m, x, allow = sc.astuple()
    ent = C.c_uint64()
    sec = lib.oracle_bench(a, gap, go, ge, m, x, allow, bases.ctypes.data, off1.ctypes.data, off2.ctypes.data,
                           len1.ctypes.data, len2.ctypes.data, len(len1), threads, C.byref(ent))
    return sec, ent.value


def expand(algo, s1, s2, start_i, start_j, end_i, end_j, ops, blank="-"):
    """ops (+ forceGlobal framing for the local algorithms, reference include/SequenceAlignment.h:156-189)
    -> (row1, row2, flags) exactly as the reference's AlignedSequence would print them."""
    a = ALGOS[algo] if isinstance(algo, str) else algo
    r1, r2, fl = [], [], []

    def up(i):
        r1.append(s1[i]); r2.append(blank); fl.append(" ")

    def left(j):
        r1.append(blank); r2.append(s2[j]); fl.append(" ")

    local = a in (1, 3)
    if local:
        for k in range(start_i):
            up(k)
        for k in range(start_j):
            left(k)
    i, j = start_i, start_j
    for op in ops:
        if op == OP_DIAG:
            r1.append(s1[i]); r2.append(s2[j]); fl.append("|" if s1[i] == s2[j] else " ")
            i += 1; j += 1
        elif op == OP_UP:
            up(i); i += 1
        else:
            left(j); j += 1
    assert (i, j) == (end_i, end_j), ((i, j), (end_i, end_j))
    if local:
        for k in range(end_i, len(s1)):
            up(k)
        for k in range(end_j, len(s2)):
            left(k)
    return "".join(r1), "".join(r2), "".join(fl)


def rescore(algo, sc, s1, s2, start_i, start_j, ops):
    """Score of an op list under the algorithm's own gap model (used for algos that expose no score:
    Hirschberg = linear gaps; MyersMiller = affine, a run of k gaps costs gap_open + k*gap_extend)."""
    a = ALGOS[algo] if isinstance(algo, str) else algo
    affine = a in (2, 3, 5)
    i, j, tot, prev = start_i, start_j, 0, -1
    for op in ops:
        if op == OP_DIAG:
            tot += sc.match if s1[i] == s2[j] else sc.mismatch
            i += 1; j += 1
        else:
            if affine:
                tot += sc.gap_extend + (sc.gap_open if op != prev else 0)
            else:
                tot += sc.gap
            if op == OP_UP:
                i += 1
            else:
                j += 1
        prev = op
    return tot
