"""Parity tests proper: the CUDA path, called through the C ABI (seqa_cuda_align_batch / seqa_ctx_*), against the
oracle on the same inputs -- bit-exact scores, end points and op strings -- plus the committed golden vectors of
the unmodified reference and size-independent properties at the BASELINE batch size."""
import os

import numpy as np
import pytest

from common import capi, check_batch_against_oracle, compare_with_oracle_batch, orc, random_pairs, scoring_to_params
from seqalib_b200 import synth
from test_oracle_golden import sc_from

pytestmark = pytest.mark.gpu
S = orc.Scoring
EDGE = [("AAAGAATGCAT", "AAACTCAT"), ("AATCG", "AACG"), ("", "ACGT"), ("ACGT", ""), ("", ""), ("A", "A"), ("A", "C"),
        ("AAAAAAAAAAAAAAAAAAAA", "AAAAA"), ("ACGTNACGT", "ACGTACGT"), ("acgt", "ACGT"), ("T" * 300, "T" * 299),
        ("AC" * 200, "CA" * 180)]

MATRIX_CASES = [
    ("nw", S.linear(-1, 2)), ("nw", S.linear(-1, 2, -1)), ("nw", S.linear(-3, 4, -2)), ("nw", S.linear(-2, 1, -1, False)),
    ("sw", S.linear(-1, 1, -1)), ("sw", S.linear(-2, 1, -1, False)), ("sw", S.linear(-1, 2)), ("sw", S.linear(-1, 2, -1)),
    ("ggotoh", S.affine(-3, -1, 1, -1)), ("ggotoh", S.affine(-3, -1, 1, -1, False)), ("ggotoh", S.affine(0, -2, 3, -1)),
    ("lgotoh", S.affine(-3, -1, 1, -1)), ("lgotoh", S.affine(-1, -1, 2, -2, False)), ("lgotoh", S.affine(-3, -1, 2, -1)),
]
LINSPACE_CASES = [
    ("hirschberg", S.linear(-1, 2, -1)), ("hirschberg", S.linear(-1, 2)), ("hirschberg", S.linear(-2, 3, -2, False)),
    ("myersmiller", S.affine(-3, -1, 1, -1)), ("myersmiller", S.affine(-3, -1, 1, -1, False)), ("myersmiller", S.affine(0, -1, 2, -1)),
]


def test_golden_vectors(gpu_lib, golden):
    groups = {}
    for v in golden:
        groups.setdefault((v["algo"], tuple(v["scoring"])), []).append(v)
    n = 0
    for (algo, sct), vs in groups.items():
        sc = sc_from(sct)
        pairs = [(v["seq1"], v["seq2"]) for v in vs]
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        res = gpu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
        for p, v in enumerate(vs):
            got = orc.expand(algo, v["seq1"], v["seq2"], int(res.start_i[p]), int(res.start_j[p]), int(res.end_i[p]),
                             int(res.end_j[p]), res.pair_ops(p))
            assert got == (v["row1"], v["row2"], v["flags"]), v
            if v["score"] is not None:
                assert int(res.score[p]) == v["score"], v
            n += 1
    assert n == len(golden)


@pytest.mark.parametrize("algo,sc", MATRIX_CASES)
def test_matrix_algorithms_random(gpu_lib, algo, sc):
    rng = np.random.default_rng(17)
    lg = algo == "lgotoh"
    dirty = [e for e in EDGE if not (lg and (not e[0] or not e[1]))]
    check_batch_against_oracle(gpu_lib, algo, sc, dirty + random_pairs(rng, 100, 1, 200))  # non-ACGT symbols -> 8-bit kernels
    pairs = [e for e in dirty if set(e[0] + e[1]) <= set("ACGT")]
    pairs += random_pairs(rng, 600, 1, 200) + random_pairs(rng, 150, 1, 120, "AC") + random_pairs(rng, 150, 1, 250, related=0.3)
    pairs += random_pairs(rng, 30, 300, 700) + random_pairs(rng, 6, 900, 1300, related=0.2)
    if lg:
        pairs = [p for p in pairs if (len(p[0]), len(p[1])) not in ((314, 288), (60, 57), (61, 58))]
    check_batch_against_oracle(gpu_lib, algo, sc, pairs)
    check_batch_against_oracle(gpu_lib, algo, sc, pairs[:400], flags=capi.FLAG_FORCE_GENERIC)
    check_batch_against_oracle(gpu_lib, algo, sc, pairs[:400], flags=capi.FLAG_TRACE8)  # 8-bit trace variants


def _kernel_used(lib, algo, sc, pairs, flags=0):
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    ctx = capi.Ctx(lib)
    ctx.upload(scoring_to_params(algo, sc, flags=flags), bases, off1, off2, len1, len2)
    ctx.run()
    ctx.sync()
    k = ctx.last_kernel()
    ctx.close()
    return k


@pytest.mark.parametrize("algo,sc", LINSPACE_CASES)
def test_linear_space_algorithms_random(gpu_lib, algo, sc):
    rng = np.random.default_rng(23)
    pairs = list(EDGE) + random_pairs(rng, 400, 1, 200) + random_pairs(rng, 100, 1, 120, "AC") + \
        random_pairs(rng, 100, 1, 250, related=0.3) + random_pairs(rng, 12, 1000, 3000) + \
        random_pairs(rng, 6, 2000, 5000, related=0.25) + [("A" * 700, "ACGT" * 150), ("ACGT" * 200, "T" * 40)]
    # a symbol outside ACGT anywhere in the batch -> the int32 sweeps (8-bit compare) take the whole batch
    check_batch_against_oracle(gpu_lib, algo, sc, pairs)
    assert _kernel_used(gpu_lib, algo, sc, pairs[:40]).endswith("_i32")
    # ACGT only -> forward + reverse sweep packed as one s16x2 wavefront
    clean = [p for p in pairs if set(p[0] + p[1]) <= set("ACGT")]
    assert _kernel_used(gpu_lib, algo, sc, clean[:40]).endswith("_s16x2")
    check_batch_against_oracle(gpu_lib, algo, sc, clean)
    check_batch_against_oracle(gpu_lib, algo, sc, clean[-24:] + clean[:100], flags=capi.FLAG_FORCE_GENERIC)
    # 32-row blocks everywhere: every sweep becomes a deep row-block pipeline across warps (progress-flag protocol)
    check_batch_against_oracle(gpu_lib, algo, sc, clean[-24:] + clean[:100], flags=capi.FLAG_LS_R1)
    check_batch_against_oracle(gpu_lib, algo, sc, pairs[-24:] + pairs[:100], flags=capi.FLAG_LS_R1)


@pytest.mark.parametrize("algo,sc,n,lo,hi", [("hirschberg", S.linear(-1, 2, -1), 6, 9000, 20000),
                                             ("myersmiller", S.affine(-3, -1, 1, -1), 6, 5000, 9000),
                                             ("hirschberg", S.linear(-4, 9, -6), 4, 6000, 9000)])
def test_linear_space_long_pairs(gpu_lib, algo, sc, n, lo, hi):
    """Sweeps tens of row blocks deep (the config-4 regime at a size the oracle finishes in seconds): unrelated and
    related pairs (10 % substitutions + indels), bit-exact including the reference's sub-optimal splits.  Scores
    leave the 16-bit range (the packed sweeps re-base per 32-column chunk); both sweep kernels are checked."""
    rng = np.random.default_rng(31)
    pairs = random_pairs(rng, n // 2, lo, hi) + random_pairs(rng, n - n // 2, lo, hi, related=0.14)
    check_batch_against_oracle(gpu_lib, algo, sc, pairs)
    check_batch_against_oracle(gpu_lib, algo, sc, pairs[:2] + pairs[-1:], flags=capi.FLAG_FORCE_GENERIC)


@pytest.mark.parametrize("algo", ["sw", "nw"])
def test_config2_shape_150bp(gpu_lib, algo):
    """a 100,000-pair sample of the BASELINE 150 bp workload (same generator and seed as
    bench.py), every field and every op bit-exact against the oracle."""
    n = 100_000
    sc = S.linear(-1, 1, -1) if algo == "sw" else S.linear(-1, 2, -1)
    ctx = capi.Ctx(gpu_lib)
    ctx.generate(scoring_to_params(algo, sc), synth.SEED, 0, n, 0, 150, 150)
    ctx.run()
    res = ctx.download(ops_capacity=n * 300)
    assert ctx.last_kernel().startswith("pk_fill")
    bases, off1, off2, l1, l2 = ctx.download_inputs(n * 300)
    hb, _, _, _, _ = synth.batch(synth.SEED, 0, 64, 0, 150, 150)
    assert np.array_equal(hb, bases[:64 * 300])  # device generator == numpy generator
    assert compare_with_oracle_batch(res, algo, sc, bases, off1, off2, l1, l2, "config2") == n


@pytest.mark.parametrize("algo,sc", [("ggotoh", S.affine(-3, -1, 1, -1)), ("lgotoh", S.affine(-3, -1, 1, -1)),
                                     ("ggotoh", S.affine(-3, -1, 1, -1, False)), ("lgotoh", S.affine(-3, -1, 1, -1, False))])
def test_config3_shape_250bp(gpu_lib, algo, sc):
    """a 100,000-pair sample of the 250 bp Gotoh (-3,-1,1,-1) workload per algorithm (+ 20,000
    pairs of the AllowMismatch=false variant) through the packed affine kernels, bit-exact against the oracle."""
    n = 100_000 if sc.allow else 20_000
    ctx = capi.Ctx(gpu_lib)
    ctx.generate(scoring_to_params(algo, sc), synth.SEED, 7_000_000, n, 0, 250, 250)
    ctx.run()
    res = ctx.download(ops_capacity=n * 500)
    assert ctx.last_kernel().startswith("pkg_fill")
    bases, off1, off2, l1, l2 = ctx.download_inputs(n * 500)
    assert compare_with_oracle_batch(res, algo, sc, bases, off1, off2, l1, l2, "config3") == n


def _rescore_all(bases, off1, off2, res, n, gap, match, mismatch):
    """Vectorised re-scoring of every alignment of a batch from its ops (linear gaps)."""
    ops_len = res.ops_len[:n].astype(np.int64)
    ops_off = res.ops_off[:n].astype(np.int64)
    tot = int(ops_len.sum())
    assert np.array_equal(ops_off, np.concatenate(([0], np.cumsum(ops_len)[:-1])))  # dense, in pair order
    ops = res.ops[:tot]
    pair = np.repeat(np.arange(n), ops_len)
    di = (ops != 2).astype(np.int64)
    dj = (ops != 1).astype(np.int64)
    ci = np.cumsum(di) - di
    cj = np.cumsum(dj) - dj
    first = ops_off[pair]
    i = ci - ci[np.minimum(first, tot - 1)] * (ops_len[pair] > 0) + res.start_i[:n].astype(np.int64)[pair]
    j = cj - cj[np.minimum(first, tot - 1)] * (ops_len[pair] > 0) + res.start_j[:n].astype(np.int64)[pair]
    a = bases[(off1[:n].astype(np.int64)[pair] + i)]
    b = bases[(off2[:n].astype(np.int64)[pair] + j)]
    s = np.where(ops == 0, np.where(a == b, match, mismatch), gap).astype(np.int64)
    score = np.zeros(n, dtype=np.int64)
    np.add.at(score, pair, s)
    ni = np.zeros(n, dtype=np.int64)
    nj = np.zeros(n, dtype=np.int64)
    np.add.at(ni, pair, di)
    np.add.at(nj, pair, dj)
    return score, ni, nj


def test_config2_full_size_properties(gpu_lib):
    """BASELINE configs[1] at full size (1,000,000 x 150 bp SW): size-independent properties of every result +
    an oracle comparison of a 3,000-pair random sample."""
    n = 1_000_000
    sc = S.linear(-1, 1, -1)
    ctx = capi.Ctx(gpu_lib)
    ctx.generate(scoring_to_params("sw", sc), synth.SEED, 0, n, 0, 150, 150)
    ctx.run()
    res = ctx.download(ops_capacity=n * 300)
    bases, off1, off2, l1, l2 = ctx.download_inputs(n * 300)
    score, ni, nj = _rescore_all(bases, off1, off2, res, n, -1, 1, -1)
    assert np.array_equal(score, res.score[:n].astype(np.int64))       # the ops re-score to the reported score
    assert np.array_equal(ni, res.end_i[:n].astype(np.int64) - res.start_i[:n])  # ops span [start,end)
    assert np.array_equal(nj, res.end_j[:n].astype(np.int64) - res.start_j[:n])
    assert (res.score[:n] > 0).all() and (res.end_i[:n] <= 150).all() and (res.end_j[:n] <= 150).all()
    # a local alignment starts and ends with a diagonal (match) step
    has = res.ops_len[:n] > 0
    assert (res.ops[res.ops_off[:n][has].astype(np.int64)] == 0).all()
    # idempotence: a second run over the resident batch gives identical results
    ctx.run()
    res2 = ctx.download(ops_capacity=n * 300)
    assert np.array_equal(res.score[:n], res2.score[:n]) and res.c.ops_used == res2.c.ops_used
    assert np.array_equal(res.ops[:res.c.ops_used], res2.ops[:res2.c.ops_used])
    rng = np.random.default_rng(99)
    for p in rng.integers(0, n, 3000):
        p = int(p)
        a = bytes(bases[int(off1[p]):int(off1[p]) + 150]).decode()
        b = bytes(bases[int(off2[p]):int(off2[p]) + 150]).decode()
        o = orc.oracle_align("sw", sc, a, b)
        assert int(res.score[p]) == o["score"] and np.array_equal(res.pair_ops(p), o["ops"]), p
        assert (int(res.start_i[p]), int(res.start_j[p]), int(res.end_i[p]), int(res.end_j[p])) == \
            (o["start_i"], o["start_j"], o["end_i"], o["end_j"]), p


def _rescore_all_affine(bases, off1, off2, res, n, go, ge, match, mismatch):
    """Vectorised re-scoring of every alignment of a batch under the affine model (a run of k gaps: go + k*ge)."""
    ops_len = res.ops_len[:n].astype(np.int64)
    ops_off = res.ops_off[:n].astype(np.int64)
    tot = int(ops_len.sum())
    ops = res.ops[:tot]
    pair = np.repeat(np.arange(n), ops_len)
    di = (ops != 2).astype(np.int64)
    dj = (ops != 1).astype(np.int64)
    ci = np.cumsum(di) - di
    cj = np.cumsum(dj) - dj
    first = ops_off[pair]
    i = ci - ci[np.minimum(first, tot - 1)] * (ops_len[pair] > 0) + res.start_i[:n].astype(np.int64)[pair]
    j = cj - cj[np.minimum(first, tot - 1)] * (ops_len[pair] > 0) + res.start_j[:n].astype(np.int64)[pair]
    a = bases[(off1[:n].astype(np.int64)[pair] + i)]
    b = bases[(off2[:n].astype(np.int64)[pair] + j)]
    prev = np.empty_like(ops)
    prev[0] = 255
    prev[1:] = ops[:-1]
    prev[ops_off[ops_len > 0]] = 255  # the first op of a pair has no predecessor
    opens = (ops != 0) & (ops != prev)
    s = np.where(ops == 0, np.where(a == b, match, mismatch), ge).astype(np.int64) + opens * go
    score = np.zeros(n, dtype=np.int64)
    np.add.at(score, pair, s)
    return score, np.bincount(pair, weights=di, minlength=n).astype(np.int64), np.bincount(pair, weights=dj, minlength=n).astype(np.int64)


@pytest.mark.parametrize("algo", ["ggotoh", "lgotoh"])
def test_config3_large_batch_properties(gpu_lib, algo):
    """BASELINE configs[2] shape at 1,000,000 x 250 bp (a tenth of the full 10 M: host memory for the numpy check):
    every alignment re-scores, under the affine model, to the score the kernel reports; ops span [start,end);
    a second run is identical; plus an oracle comparison of a 1,500-pair random sample."""
    n = 1_000_000
    sc = S.affine(-3, -1, 1, -1)
    ctx = capi.Ctx(gpu_lib)
    ctx.generate(scoring_to_params(algo, sc), synth.SEED, 20_000_000, n, 0, 250, 250)
    ctx.run()
    res = ctx.download(ops_capacity=n * 500)
    assert ctx.last_kernel().startswith("pkg_fill")
    bases, off1, off2, l1, l2 = ctx.download_inputs(n * 500)
    score, ni, nj = _rescore_all_affine(bases, off1, off2, res, n, -3, -1, 1, -1)
    assert np.array_equal(score, res.score[:n].astype(np.int64))
    assert np.array_equal(ni, res.end_i[:n].astype(np.int64) - res.start_i[:n])
    assert np.array_equal(nj, res.end_j[:n].astype(np.int64) - res.start_j[:n])
    if algo == "ggotoh":
        assert (res.start_i[:n] == 0).all() and (res.end_i[:n] == 250).all() and (res.end_j[:n] == 250).all()
    ctx.run()
    res2 = ctx.download(ops_capacity=n * 500)
    assert np.array_equal(res.score[:n], res2.score[:n]) and res.c.ops_used == res2.c.ops_used
    assert np.array_equal(res.ops[:res.c.ops_used], res2.ops[:res2.c.ops_used])
    rng = np.random.default_rng(7)
    for p in rng.integers(0, n, 1500):
        p = int(p)
        a = bytes(bases[int(off1[p]):int(off1[p]) + 250]).decode()
        b = bytes(bases[int(off2[p]):int(off2[p]) + 250]).decode()
        o = orc.oracle_align(algo, sc, a, b)
        assert int(res.score[p]) == o["score"] and np.array_equal(res.pair_ops(p), o["ops"]), p


def test_config4_full_length_pairs(gpu_lib):
    """BASELINE configs[3] at its real pair length: 8 Hirschberg pairs of 100,000 x 100,000 bp (random DNA from the
    shared generator) and 4 MyersMiller pairs of 30,000 bp, bit-exact against the oracle (one pair per host thread)."""
    from concurrent.futures import ThreadPoolExecutor
    for algo, sc, n, L in (("hirschberg", S.linear(-1, 2, -1), 8, 100_000), ("myersmiller", S.affine(-3, -1, 1, -1), 4, 30_000)):
        ctx = capi.Ctx(gpu_lib)
        ctx.generate(scoring_to_params(algo, sc), synth.SEED, 0, n, 0, L, L)
        ctx.run()
        res = ctx.download(ops_capacity=n * 2 * L)
        assert ctx.last_kernel().endswith("_s16x2")
        bases, off1, off2, l1, l2 = ctx.download_inputs(n * 2 * L)
        seqs = [(bytes(bases[int(off1[p]):int(off1[p]) + L]).decode(), bytes(bases[int(off2[p]):int(off2[p]) + L]).decode())
                for p in range(n)]
        with ThreadPoolExecutor(max_workers=n) as ex:
            outs = list(ex.map(lambda ab: orc.oracle_align(algo, sc, ab[0], ab[1]), seqs))
        for p, o in enumerate(outs):
            assert int(res.score[p]) == o["score"], (algo, p)
            assert np.array_equal(res.pair_ops(p), o["ops"]), (algo, p)


def test_config4_related_pairs(gpu_lib):
    """BASELINE configs[3], realism variant: RELATED pairs -- sequence 2 = sequence 1 with 10 %
    substitutions, 2 % insertions, 2 % deletions from the shared generator (seqalib_b200/synth.py: related_sequence
    documents the exact procedure).  Long diagonal runs, a different split pattern, scores far beyond 16 bits
    (the packed sweeps re-base): 8 Hirschberg pairs of 100,000 bp and 8 MyersMiller pairs of 30,000 bp, bit-exact."""
    for algo, sc, n, L in (("hirschberg", S.linear(-1, 2, -1), 8, 100_000), ("myersmiller", S.affine(-3, -1, 1, -1), 8, 30_000)):
        bases, off1, off2, l1, l2 = synth.related_batch(synth.SEED, 0, n, L)
        ctx = capi.Ctx(gpu_lib)
        ctx.upload(scoring_to_params(algo, sc), bases, off1, off2, l1, l2)
        ctx.run()
        res = ctx.download(ops_capacity=int(l1.sum() + l2.sum()))
        assert ctx.last_kernel().endswith("_s16x2")
        ctx.close()
        assert compare_with_oracle_batch(res, algo, sc, bases, off1, off2, l1, l2, "config4 related") == n
        ident = [(res.pair_ops(p) == 0).sum() / float(l1[p]) for p in range(n)]
        assert min(ident) > 0.9  # the pairs really are related: > 90 % of sequence 1 sits on diagonal steps


def test_two_bit_wire_format(gpu_lib):
    """SEQA_FLAG_OPS_2BIT (4 ops per byte over PCIe): same alignments as the one-byte form, through the one-shot call
    in several waves (300,000 pairs > one wave) and through the resident interface; sampled against the oracle."""
    n = 300_000
    sc = S.linear(-1, 1, -1)
    bases, off1, off2, l1, l2 = None, None, None, None, None
    ctx = capi.Ctx(gpu_lib)
    ctx.generate(scoring_to_params("sw", sc, flags=capi.FLAG_OPS_2BIT), synth.SEED, 11_000_000, n, 0, 150, 150)
    ctx.run()
    r2 = ctx.download(ops_capacity=n * 76)
    bases, off1, off2, l1, l2 = ctx.download_inputs(n * 300)
    ctx.close()
    assert r2.packed2 and int(r2.ops_off[-1]) < n * 76
    plain = gpu_lib.align_batch(scoring_to_params("sw", sc), bases, off1, off2, l1, l2)
    packed = gpu_lib.align_batch(scoring_to_params("sw", sc, flags=capi.FLAG_OPS_2BIT), bases, off1, off2, l1, l2,
                                 capi.Results(n, n * 76))
    for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
        assert np.array_equal(getattr(plain, name), getattr(packed, name)), name
        assert np.array_equal(getattr(plain, name), getattr(r2, name)), name
    rng = np.random.default_rng(2)
    for p in [0, 1, n - 1] + [int(x) for x in rng.integers(0, n, 3000)]:
        assert np.array_equal(plain.pair_ops(p), packed.pair_ops(p)), p
        assert np.array_equal(plain.pair_ops(p), r2.pair_ops(p)), p
    idx = np.sort(rng.choice(n, 2000, replace=False))
    for p in idx:
        a = bytes(bases[int(off1[p]):int(off1[p]) + 150]).decode()
        b = bytes(bases[int(off2[p]):int(off2[p]) + 150]).decode()
        assert np.array_equal(packed.pair_ops(p), orc.oracle_align("sw", sc, a, b)["ops"]), p


def test_config4_myersmiller_full_length(gpu_lib):
    """BASELINE configs[3] at its real pair length for MyersMiller too: 8 pairs of 100,000 x 100,000 bp (4 random, 4
    related), bit-exact against the oracle (about a minute: one pair per host thread)."""
    sc = S.affine(-3, -1, 1, -1)
    L = 100_000
    rb, ro1, ro2, rl1, rl2 = synth.batch(synth.SEED, 100, 4, 0, L, L)
    qb, qo1, qo2, ql1, ql2 = synth.related_batch(synth.SEED, 200, 4, L)
    bases = np.concatenate([rb, qb])
    off1 = np.concatenate([ro1, qo1 + np.uint64(len(rb))])
    off2 = np.concatenate([ro2, qo2 + np.uint64(len(rb))])
    l1, l2 = np.concatenate([rl1, ql1]), np.concatenate([rl2, ql2])
    ctx = capi.Ctx(gpu_lib)
    ctx.upload(scoring_to_params("myersmiller", sc), bases, off1, off2, l1, l2)
    ctx.run()
    res = ctx.download(ops_capacity=int(l1.sum() + l2.sum()))
    assert ctx.last_kernel().endswith("_s16x2")
    ctx.close()
    assert compare_with_oracle_batch(res, "myersmiller", sc, bases, off1, off2, l1, l2, "config4 MM 100 kbp") == 8


def test_mixed_length_batch(gpu_lib):
    """a 100,000-pair sample of the mixed-length workload (independent U[50,1000] lengths,
    NW (-1,2,-1) and SW (-1,1,-1) over the same pairs, generated on the device and read back for the oracle)."""
    n = 100_000
    for algo, sc in (("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1))):
        ctx = capi.Ctx(gpu_lib)
        ctx.generate(scoring_to_params(algo, sc), synth.SEED, 5_000_000, n, 1)
        ctx.run()
        l1, l2 = synth.lengths(synth.SEED, 5_000_000, n)
        tot = int(l1.sum() + l2.sum())
        res = ctx.download(ops_capacity=tot)
        assert ctx.last_kernel().startswith("pk_fill")
        bases, off1, off2, d1, d2 = ctx.download_inputs(tot)
        assert np.array_equal(d1, l1) and np.array_equal(d2, l2)
        assert compare_with_oracle_batch(res, algo, sc, bases, off1, off2, l1, l2, "config5") == n
        ctx.close()


def test_multi_device_split(gpu_lib):
    if gpu_lib.device_count() < 2:
        pytest.skip("one device visible")
    rng = np.random.default_rng(5)
    pairs = random_pairs(rng, 500, 1, 200)
    check_batch_against_oracle(gpu_lib, "sw", S.linear(-1, 1, -1), pairs, device_count=2)
    check_batch_against_oracle(gpu_lib, "ggotoh", S.affine(-3, -1, 1, -1), pairs, device_count=gpu_lib.device_count())
    # the 2-bit wire format places every wave behind ceil(len/4) bytes per earlier pair, on whichever device it ran
    check_batch_against_oracle(gpu_lib, "nw", S.linear(-1, 2, -1), pairs, flags=capi.FLAG_OPS_2BIT, device_count=2)


def test_device_resident_results(gpu_lib):
    """seqa_ctx_device_results: the arrays a GPU consumer would read equal what seqa_ctx_download copies out."""
    import ctypes as C
    import glob
    import site
    cands = ["/usr/local/cuda/lib64/libcudart.so"]
    for sp in site.getsitepackages():
        cands += glob.glob(os.path.join(sp, "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = C.CDLL([c for c in cands if os.path.exists(c)][-1])
    n = 5000
    ctx = capi.Ctx(gpu_lib)
    ctx.generate(scoring_to_params("sw", S.linear(-1, 1, -1)), synth.SEED, 123, n, 0, 150, 150)
    ctx.run()
    dev = ctx.device_results()
    res = ctx.download(ops_capacity=n * 300)
    assert int(dev.ops_used) == int(res.c.ops_used) > 0

    def d2h(ptr, nbytes, dtype):
        out = np.zeros(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        addr = int(ptr)
        assert rt.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(addr), C.c_size_t(nbytes), 2) == 0
        return out
    assert np.array_equal(d2h(dev.score, n * 4, np.int32), res.score[:n])
    assert np.array_equal(d2h(dev.ops_len, n * 4, np.uint32), res.ops_len[:n])
    assert np.array_equal(d2h(dev.ops_off, n * 8, np.uint64), res.ops_off[:n])
    assert np.array_equal(d2h(dev.ops, int(dev.ops_used), np.uint8), res.ops[:int(res.c.ops_used)])


def test_int_peak_microbenchmark(gpu_lib):
    import ctypes as C
    v, mhz = C.c_double(), C.c_double()
    gpu_lib.check(gpu_lib.L.seqa_cuda_int_peak(0, 0, C.byref(v), C.byref(mhz)))
    assert 16 <= v.value <= 140 and 500 <= mhz.value <= 2500
