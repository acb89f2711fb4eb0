"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as orc  # noqa: E402
from seqalib_b200 import capi  # noqa: E402

EMU_SO = os.path.join(ROOT, "tests", "emu", "libseqa_emu.so")


def scoring_to_params(algo, sc, **kw):
    """oracle Scoring -> C-ABI seqa_params (allow=False reproduces the 2-argument ScoringSystem)."""
    return capi.make_params(algo, gap=sc.gap, gap_open=sc.gap_open, gap_extend=sc.gap_extend, match=sc.match,
                            mismatch=sc.mismatch if sc.allow else 0, allow=sc.allow, **kw)


def random_pairs(rng, n, lo, hi, alphabet="ACGT", related=0.0):
    out = []
    for _ in range(n):
        l1, l2 = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        a = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), l1))
        if related > 0 and l1 > 0:
            b = []
            for ch in a:
                r = rng.random()
                if r < related * 0.6:
                    b.append(alphabet[int(rng.integers(0, len(alphabet)))])
                elif r < related * 0.8:
                    continue
                elif r < related:
                    b.append(ch)
                    b.append(alphabet[int(rng.integers(0, len(alphabet)))])
                else:
                    b.append(ch)
            b = "".join(b)[:max(hi, 1)]
            if lo >= 1 and not b:
                b = alphabet[0]
        else:
            b = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), l2))
        out.append((a, b))
    return out


def check_batch_against_oracle(lib, algo, sc, pairs, flags=0, device_count=1, label=""):
    """Run `pairs` through seqa_cuda_align_batch and compare every field with the C oracle. Returns #pairs."""
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    prm = scoring_to_params(algo, sc, flags=flags, device_count=device_count)
    res = lib.align_batch(prm, bases, off1, off2, len1, len2)
    for p, (a, b) in enumerate(pairs):
        o = orc.oracle_align(algo, sc, a, b)
        got = dict(score=int(res.score[p]), start_i=int(res.start_i[p]), start_j=int(res.start_j[p]),
                   end_i=int(res.end_i[p]), end_j=int(res.end_j[p]))
        exp = {k: o[k] for k in got}
        ops = res.pair_ops(p)
        if got != exp or not np.array_equal(ops, o["ops"]):
            raise AssertionError("%s %s %r pair %d (%s | %s): got %r ops %s, expected %r ops %s" % (
                label, algo, sc, p, a, b, got, ops.tolist(), exp, o["ops"].tolist()))
    return len(pairs)


def _ragged_equal(ops_a, off_a, ops_b, off_b, lens):
    """Index of the first pair whose op strings differ (ops_x[off_x[p] : off_x[p]+lens[p]]), or -1.  Vectorised."""
    lens = lens.astype(np.int64)
    tot = int(lens.sum())
    if tot == 0:
        return -1
    owner = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    start = np.cumsum(lens) - lens
    within = np.arange(tot, dtype=np.int64) - start[owner]
    a = ops_a[off_a.astype(np.int64)[owner] + within]
    b = ops_b[off_b.astype(np.int64)[owner] + within]
    bad = np.nonzero(a != b)[0]
    return int(owner[bad[0]]) if len(bad) else -1


def compare_with_oracle_batch(res, algo, sc, bases, off1, off2, len1, len2, label=""):
    """Every field of every pair of a GPU result (capi.Results) against the threaded C oracle, bit-exact.
    Used for the >= 100 k-pair parity samples of SURVEY.md 8d.  Returns the number of pairs compared."""
    o = orc.oracle_align_batch(algo, sc, bases, off1, off2, len1, len2)
    n = len(len1)
    for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
        g, e = getattr(res, name)[:n], getattr(o, name)[:n]
        bad = np.nonzero(g != e)[0]
        assert len(bad) == 0, "%s %s %r: %s of pair %d is %d, oracle %d (%d pairs differ)" % (
            label, algo, sc, name, bad[0], g[bad[0]], e[bad[0]], len(bad))
    p = _ragged_equal(res.ops, res.ops_off[:n], o.ops, o.slot_off, o.ops_len)
    assert p < 0, "%s %s %r: ops of pair %d differ: got %s, oracle %s" % (
        label, algo, sc, p, res.pair_ops(p).tolist(), o.pair_ops(p).tolist())
    return n


def prep_staging_layouts(rng, n_pairs, length):
    """Uniform pairs in three layouts: dense; every sequence 1,000 bytes from the next (a job's span outgrows the staging
    buffer of pk_prep_kernel: global loads); the first 64 pairs dense and the rest spread out (both forms in one launch)."""
    pairs = random_pairs(rng, n_pairs, length, length)
    out = [orc.batch_arrays(pairs) + (pairs,)]
    for dense_head in (0, 64):
        chunks, o1, o2, pos = [], np.zeros(n_pairs, np.uint64), np.zeros(n_pairs, np.uint64), 3
        chunks.append(b"NNN")
        for p, (a, b) in enumerate(pairs):
            gap = b"" if p < dense_head else b"N" * 1000
            o1[p] = pos
            chunks += [a.encode(), gap]
            pos += len(a) + len(gap)
            o2[p] = pos
            chunks += [b.encode(), gap]
            pos += len(b) + len(gap)
        l = np.full(n_pairs, length, np.uint32)
        out.append((np.frombuffer(b"".join(chunks), dtype=np.uint8).copy(), o1, o2, l, l.copy(), pairs))
    return out
