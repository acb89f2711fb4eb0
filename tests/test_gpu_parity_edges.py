"""Parity holes closed in round 2 (VERDICT r1 "What's weak" #1, "What's missing" #6): the CUDA path directly against
the compiled UNMODIFIED reference (oracle/_ref), scorings and shapes on both sides of every host-side threshold that
selects a kernel variant, long pairs through the full-matrix aligners inside mixed batches, SEQA_FLAG_SCORE_ONLY, and
the 1,000-pair 1-20 kbp sample of SURVEY.md 8d config 4.  All through the C ABI, bit-exact."""
import json
import os
import zlib

import numpy as np
import pytest

from common import ROOT, capi, check_batch_against_oracle, compare_with_oracle_batch, orc, prep_staging_layouts, random_pairs, scoring_to_params
from seqalib_b200 import synth

pytestmark = pytest.mark.gpu
S = orc.Scoring


def _seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[k] for k in rng.integers(0, len(alphabet), n))


UB_SHAPES = ((314, 288), (60, 57), (61, 58))  # LocalGotoh: undefined behaviour in the reference, rejected per pair


def _defined(algo, pairs):
    """drop what the reference leaves undefined for LocalGotoh (the three UB shapes, empty inputs)"""
    if algo != "lgotoh":
        return pairs
    return [p for p in pairs if (len(p[0]), len(p[1])) not in UB_SHAPES and p[0] and p[1]]


def _kernel_used(lib, algo, sc, pairs, flags=0):
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    ctx = capi.Ctx(lib)
    ctx.upload(scoring_to_params(algo, sc, flags=flags), bases, off1, off2, len1, len2)
    ctx.run()
    ctx.sync()
    k = ctx.last_kernel()
    ctx.close()
    return k


# ---- (a) CUDA vs the compiled reference itself ---------------------------------------------------------------------
REF_CASES = [("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1)), ("ggotoh", S.affine(-3, -1, 1, -1)),
             ("lgotoh", S.affine(-3, -1, 1, -1)), ("hirschberg", S.linear(-1, 2, -1)), ("myersmiller", S.affine(-3, -1, 1, -1))]


@pytest.mark.parametrize("algo,sc", REF_CASES)
def test_cuda_against_compiled_reference(gpu_lib, algo, sc):
    """2,000 pairs per aligner: the rows the UNMODIFIED reference prints (oracle/_ref/libseqa_ref.so, built from
    /root/reference/include by oracle/Makefile and shipped to the GPU box) against the rows expanded from the CUDA ops
    (forceGlobal framing included), plus the score where the reference exposes one.  No C restatement in between."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libseqa_ref.so was not prebuilt (needs /root/reference at build time)")
    rng = np.random.default_rng(2024)
    pairs = random_pairs(rng, 1200, 1, 200) + random_pairs(rng, 300, 1, 120, "AC") + random_pairs(rng, 500, 20, 250, related=0.3)
    pairs = _defined(algo, pairs)
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    res = gpu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
    for p, (a, b) in enumerate(pairs):
        r = orc.ref_align(algo, sc, a, b, functor=(algo == "sw"))  # SW is only defined with a functor (SASmithWaterman.h:53-54)
        got = orc.expand(algo, a, b, int(res.start_i[p]), int(res.start_j[p]), int(res.end_i[p]), int(res.end_j[p]), res.pair_ops(p))
        assert got == (r["row1"], r["row2"], r["flags"]), (algo, p, a, b)
        if r["score"] is not None and algo in ("nw", "sw", "ggotoh", "lgotoh"):
            assert int(res.score[p]) == r["score"], (algo, p)


# ---- (b) both sides of every host threshold ------------------------------------------------------------------------
def test_trace_width_threshold_linear(gpu_lib):
    """packed_trace_bits: 4 trace bits iff Match + |Mismatch| + 2|Gap| <= 7 (seqa_cuda.cu)."""
    rng = np.random.default_rng(1)
    pairs = random_pairs(rng, 300, 1, 200) + random_pairs(rng, 80, 1, 150, "AC") + random_pairs(rng, 80, 50, 250, related=0.3)
    for algo in ("nw", "sw"):
        for sc, want in ((S.linear(-2, 2, -1), "_t4"), (S.linear(-2, 2, -2), "_t8"), (S.linear(-1, 4, -1), "_t4"), (S.linear(-1, 4, -2), "_t8"),
                         (S.linear(-3, 1, -1, False), "_t4"), (S.linear(-3, 2, -1, False), "_t8")):
            assert _kernel_used(gpu_lib, algo, sc, pairs[:70]).endswith(want), (algo, sc)
            check_batch_against_oracle(gpu_lib, algo, sc, pairs, label="tb-threshold")


def test_trace_width_threshold_affine(gpu_lib):
    """packed_affine_trace_bits: 4 bits iff m + x + 2g <= 12 and m + g <= 7, g = -(GapOpen + GapExtend)."""
    rng = np.random.default_rng(2)
    pairs = random_pairs(rng, 250, 1, 200) + random_pairs(rng, 60, 1, 150, "AC") + random_pairs(rng, 60, 50, 250, related=0.3)
    for algo in ("ggotoh", "lgotoh"):
        for sc, want in ((S.affine(-2, -1, 4, -2), "_t4"), (S.affine(-2, -1, 4, -3), "_t8"), (S.affine(-3, -1, 3, -1), "_t4"),
                         (S.affine(-3, -1, 4, -1), "_t8"), (S.affine(0, -1, 6, -4), "_t4"), (S.affine(0, -1, 7, -1), "_t8")):
            assert _kernel_used(gpu_lib, algo, sc, pairs[:70]).endswith(want), (algo, sc)
            check_batch_against_oracle(gpu_lib, algo, sc, _defined(algo, pairs), label="affine-tb-threshold")


def test_packed_versus_generic_scoring_threshold(gpu_lib):
    """packed_scoring_ok: m + x + 2g <= 120 (linear) / m + x + 3g - GapOpen <= 120 (affine) keeps the s16x2 kernels."""
    rng = np.random.default_rng(3)
    short = random_pairs(rng, 200, 1, 45) + random_pairs(rng, 50, 1, 40, "AC")
    for algo, sc, packed in (("nw", S.linear(-20, 50, -30), True), ("nw", S.linear(-20, 50, -31), False),
                             ("sw", S.linear(-20, 50, -30), True), ("sw", S.linear(-20, 50, -31), False),
                             ("ggotoh", S.affine(-10, -10, 30, -20), True), ("ggotoh", S.affine(-10, -10, 30, -21), False),
                             ("lgotoh", S.affine(-10, -10, 30, -20), True), ("lgotoh", S.affine(-10, -10, 30, -21), False),
                             ("nw", S.linear(-50, 1, -1), True), ("nw", S.linear(-51, 1, -1), False),
                             ("sw", S.linear(-1, 100, -1), True), ("sw", S.linear(-1, 101, -1), False)):
        k = _kernel_used(gpu_lib, algo, sc, short[:70])
        assert k.startswith("pk") == packed, (algo, sc, k)
        check_batch_against_oracle(gpu_lib, algo, sc, _defined(algo, short), label="packed-threshold")


def test_packed_score_range_limit(gpu_lib):
    """packed_shape_ok: a pair leaves the 16-bit kernels when (M+N+34)|Gap| + 300 or min(M,N) Match + 300 reaches 30,000
    (affine: (M+N+34) x unit reaches 9,000); batches that hold pairs on BOTH sides run packed and generic kernels together."""
    rng = np.random.default_rng(4)
    filler = random_pairs(rng, 130, 100, 200)
    both = [(_seq(rng, 279), _seq(rng, 280)), (_seq(rng, 280), _seq(rng, 280)), (_seq(rng, 281), _seq(rng, 279))]
    check_batch_against_oracle(gpu_lib, "nw", S.linear(-50, 1, -1), both + filler, label="lo-limit")
    check_batch_against_oracle(gpu_lib, "sw", S.linear(-50, 1, -1), both + filler, label="lo-limit")
    tall = [(_seq(rng, 296, "A"), _seq(rng, 296, "A")), (_seq(rng, 297, "A"), _seq(rng, 297, "A")), (_seq(rng, 297, "AC"), _seq(rng, 296, "AC"))]
    check_batch_against_oracle(gpu_lib, "nw", S.linear(-1, 100, -1), tall + filler, label="hi-limit")
    check_batch_against_oracle(gpu_lib, "sw", S.linear(-1, 100, -1), tall + filler, label="hi-limit")
    aff = [(_seq(rng, 733), _seq(rng, 732)), (_seq(rng, 733), _seq(rng, 733)), (_seq(rng, 700), _seq(rng, 766))] + random_pairs(rng, 3, 720, 740, related=0.2)
    for algo in ("ggotoh", "lgotoh"):
        check_batch_against_oracle(gpu_lib, algo, S.affine(-3, -1, 1, -1), _defined(algo, aff + filler), label="affine-range-limit")


# ---- (c) shared-memory / global strip boundary, packed length cap, long pairs in mixed batches ---------------------
def test_strip_boundary_storage_switch_320_321(gpu_lib):
    """PK_MAX_LEN: linear-gap pairs up to 320 columns keep the strip boundary in shared memory; one pair of 321 columns
    moves the batch to the global-row variant."""
    rng = np.random.default_rng(5)
    base = random_pairs(rng, 120, 100, 320) + [(_seq(rng, 320), _seq(rng, 320)), (_seq(rng, 17), _seq(rng, 320)), (_seq(rng, 320), _seq(rng, 33))]
    more = base + [(_seq(rng, 320), _seq(rng, 321)), (_seq(rng, 321), _seq(rng, 320)), (_seq(rng, 5), _seq(rng, 321))]
    for algo, sc in (("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1)), ("sw", S.linear(-2, 3, -2))):
        check_batch_against_oracle(gpu_lib, algo, sc, base, label="<=320")
        check_batch_against_oracle(gpu_lib, algo, sc, more, label="321")


def test_strip_boundary_rows_as_8_bit_differences(gpu_lib):
    """The strip-boundary rows in global memory travel as 8-bit fields (differences along the row; affine: + H - Ix): scorings
    that push the differences to both ends of their windows, identical / disjoint / related long pairs, AllowMismatch off."""
    rng = np.random.default_rng(29)
    a = _seq(rng, 330)
    pairs = random_pairs(rng, 60, 321, 420) + random_pairs(rng, 20, 330, 400, related=0.15) + random_pairs(rng, 20, 40, 350, "AC") + \
        [(a, a), ("ACGT" * 90, "ACGT" * 90), ("A" * 340, "C" * 330), ("A" * 25, "ACGT" * 85), ("ACGT" * 85, "A" * 25), (a, a[::-1])]
    for algo, sc in (("nw", S.linear(-14, 75, -1)), ("nw", S.linear(-1, 70, -40)), ("sw", S.linear(-29, 30, -30)), ("sw", S.linear(-1, 1, -1)),
                     ("nw", S.linear(-2, 3, -1, False)), ("sw", S.linear(-50, 19, -1))):
        check_batch_against_oracle(gpu_lib, algo, sc, pairs, label="gb-rows")
    short = [(x[:260], y[:250]) for x, y in pairs]
    for algo, sc in (("ggotoh", S.affine(-3, -1, 1, -1)), ("ggotoh", S.affine(-10, -40, 20, -30)), ("lgotoh", S.affine(-5, -5, 12, -9)),
                     ("ggotoh", S.affine(0, -1, 7, -5)), ("lgotoh", S.affine(-3, -1, 1, -1, False)), ("ggotoh", S.affine(-40, -10, 30, -1))):
        check_batch_against_oracle(gpu_lib, algo, sc, _defined(algo, short), label="affine-rows")


def test_packed_length_cap_2048_2049(gpu_lib):
    """PKG_MAX_LEN: a side of 2,048 still runs on the s16x2 kernels, 2,049 goes to the int32 wavefront -- in one batch."""
    rng = np.random.default_rng(6)
    pairs = [(_seq(rng, 2048), _seq(rng, 2048)), (_seq(rng, 2049), _seq(rng, 2048)), (_seq(rng, 2048), _seq(rng, 2049)),
             (_seq(rng, 40), _seq(rng, 2048)), (_seq(rng, 2049), _seq(rng, 40))] + random_pairs(rng, 2, 2000, 2048, related=0.2)
    pairs += random_pairs(rng, 130, 100, 200)
    for algo, sc in (("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1))):
        check_batch_against_oracle(gpu_lib, algo, sc, pairs, label="2048|2049")


@pytest.mark.parametrize("algo,sc,L", [("nw", S.linear(-1, 2, -1), 12000), ("sw", S.linear(-1, 1, -1), 12000),
                                       ("ggotoh", S.affine(-3, -1, 1, -1), 7000), ("lgotoh", S.affine(-3, -1, 1, -1), 7000)])
def test_long_pairs_full_matrix_in_mixed_batch(gpu_lib, algo, sc, L):
    """A 5-20 kbp pair through each full-matrix aligner (the int32 warp wavefront crossing ~50-90 row blocks of 128
    rows, boundary rows in global memory) in a batch that also holds 150 bp pairs on the packed kernels; unrelated and
    related (long diagonal runs) variants."""
    rng = np.random.default_rng(7)
    long_pairs = [(_seq(rng, L), _seq(rng, L - 777))] + random_pairs(rng, 1, L // 2, L // 2 + 500, related=0.15)
    pairs = random_pairs(rng, 70, 150, 150) + long_pairs + random_pairs(rng, 70, 150, 150)
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    res = gpu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
    assert compare_with_oracle_batch(res, algo, sc, bases, off1, off2, len1, len2, "long+short") == len(pairs)


def test_batch_mixing_packed_eligible_and_ineligible_pairs(gpu_lib):
    """Empty sequences, 16-bit-range overflows and > 2,048 bp pairs (generic kernels) interleaved with packed pairs."""
    rng = np.random.default_rng(8)
    pairs = []
    for k in range(150):
        pairs.append((_seq(rng, int(rng.integers(1, 300))), _seq(rng, int(rng.integers(1, 300)))))
        if k % 10 == 0:
            pairs.append(("", _seq(rng, 30)))
        if k % 25 == 0:
            pairs.append((_seq(rng, 2100), _seq(rng, 50)))
        if k % 40 == 0:
            pairs.append((_seq(rng, 900, "A"), _seq(rng, 900, "A")))
    for algo, sc in (("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 40, -1)), ("ggotoh", S.affine(-3, -1, 1, -1))):
        check_batch_against_oracle(gpu_lib, algo, sc, pairs, label="mixed eligibility")
        check_batch_against_oracle(gpu_lib, algo, sc, pairs, flags=capi.FLAG_OPS_2BIT, label="mixed eligibility 2-bit")


# ---- (d) SEQA_FLAG_SCORE_ONLY --------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo,sc", [("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1)), ("ggotoh", S.affine(-3, -1, 1, -1)),
                                     ("lgotoh", S.affine(-3, -1, 1, -1))])
def test_score_only_flag(gpu_lib, algo, sc):
    """SEQA_FLAG_SCORE_ONLY: scores and end cells of the DP fill without any traceback, against the oracle; the start /
    ops arrays may be NULL; packed and generic kernels."""
    import ctypes as C
    rng = np.random.default_rng(9)
    pairs = _defined(algo, random_pairs(rng, 400, 1, 250) + random_pairs(rng, 40, 1, 100, "AC"))
    for flags in (capi.FLAG_SCORE_ONLY, capi.FLAG_SCORE_ONLY | capi.FLAG_FORCE_GENERIC):
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        n = len(pairs)
        score = np.zeros(n, np.int32)
        end_i, end_j = np.full(n, 7777, np.uint32), np.full(n, 7777, np.uint32)
        out = capi.BatchOut(score.ctypes.data, None, None, end_i.ctypes.data, end_j.ctypes.data, None, None, None, 0, 0)
        bi = capi.Lib.batch_in(bases, off1, off2, len1, len2)
        prm = scoring_to_params(algo, sc, flags=flags)
        gpu_lib.check(gpu_lib.L.seqa_cuda_align_batch(C.byref(prm), C.byref(bi), C.byref(out)))
        local = algo in ("sw", "lgotoh")
        for p, (a, b) in enumerate(pairs):
            o = orc.oracle_align(algo, sc, a, b)
            assert int(score[p]) == o["score"], (algo, flags, p)
            assert int(end_i[p]) == o["end_i"], (algo, flags, p)  # global: M; local: MaxRow
            if not local:
                assert int(end_j[p]) == o["end_j"], (algo, flags, p)  # (local MaxCol is resolved by the walk, which is skipped)
        assert out.ops_used == 0


# ---- (e) SURVEY 8d config 4: 1,000 pairs of 1-20 kbp per linear-space aligner --------------------------------------
@pytest.mark.parametrize("algo,sc", [("hirschberg", S.linear(-1, 2, -1)), ("myersmiller", S.affine(-3, -1, 1, -1))])
def test_config4_thousand_pairs_1_to_20_kbp(gpu_lib, algo, sc):
    """1,000 pairs with independent U[1000, 20000] lengths (half unrelated, half related: the shared generator's
    sequence 1 against a mutated copy of a prefix of it), every op bit-exact against the threaded C oracle."""
    rng = np.random.default_rng(10)
    n = 1000
    l1 = rng.integers(1000, 20001, n).astype(np.uint32)
    l2 = rng.integers(1000, 20001, n).astype(np.uint32)
    seqs = []
    for p in range(n):
        a = synth.sequence(synth.SEED, 900_000 + p, 0, int(l1[p]))
        if p % 2:
            _, b = synth.related_sequence(synth.SEED, 900_000 + p, int(l1[p]))
            b = b[:int(l2[p])]
        else:
            b = synth.sequence(synth.SEED, 900_000 + p, 1, int(l2[p]))
        seqs.append((a, b))
    l1 = np.array([len(a) for a, _ in seqs], np.uint32)
    l2 = np.array([len(b) for _, b in seqs], np.uint32)
    tot = l1.astype(np.uint64) + l2
    off1 = np.zeros(n, np.uint64)
    off1[1:] = np.cumsum(tot)[:-1]
    off2 = off1 + l1
    bases = np.concatenate([np.concatenate(ab) for ab in seqs])
    ctx = capi.Ctx(gpu_lib)
    ctx.upload(scoring_to_params(algo, sc), bases, off1, off2, l1, l2)
    ctx.run()
    res = ctx.download(ops_capacity=int(tot.sum()))
    assert ctx.last_kernel().endswith("_s16x2")
    ctx.close()
    assert compare_with_oracle_batch(res, algo, sc, bases, off1, off2, l1, l2, "config4 1-20 kbp") == n


def test_config4_full_size_fingerprints(gpu_lib):
    """BASELINE configs[3] at FULL size, all pairs: 64 x 100 kbp random DNA through HirschbergSA and MyersMillerSA,
    every pair's (score, ops_len, CRC-32 of the op string) against tests/golden/config4_100kbp.json -- fingerprints of
    the C oracle's output on the same generator pairs, produced offline by oracle/make_golden_config4.py."""
    path = os.path.join(ROOT, "tests", "golden", "config4_100kbp.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/config4_100kbp.json not generated yet")
    gold = json.load(open(path))
    n, L = gold["pairs"], gold["len"]
    for algo, sc in (("hirschberg", S.linear(-1, 2, -1)), ("myersmiller", S.affine(-3, -1, 1, -1))):
        if algo not in gold["algos"]:
            continue
        assert gold["algos"][algo]["scoring"] == list(sc.astuple())
        ctx = capi.Ctx(gpu_lib)
        ctx.generate(scoring_to_params(algo, sc), gold["seed"], gold["first_pair"], n, 0, L, L)
        ctx.run()
        res = ctx.download(ops_capacity=n * 2 * L)
        ctx.close()
        for p, g in enumerate(gold["algos"][algo]["pairs"]):
            ops = np.ascontiguousarray(res.pair_ops(p))
            assert (int(res.score[p]), int(res.ops_len[p]), zlib.crc32(ops.tobytes()) & 0xffffffff) == (g["score"], g["ops_len"], g["crc32"]), (algo, p)


def test_local_gotoh_undefined_shapes_per_pair(gpu_lib):
    """One LocalGotoh pair of shape (60,57) / (61,58) / (314,288) (undefined behaviour in the reference,
    SALocalGotoh.h:484-488) no longer fails the batch: it is reported with ops_len = SEQA_PAIR_UNSUPPORTED."""
    rng = np.random.default_rng(11)
    sc = S.affine(-3, -1, 1, -1)
    pairs = random_pairs(rng, 300, 20, 200)
    ub = {17: (60, 57), 130: (61, 58), 222: (314, 288)}
    for k, (m, n_) in ub.items():
        pairs[k] = (_seq(rng, m), _seq(rng, n_))
    pairs = [p if (k in ub or (len(p[0]), len(p[1])) not in ub.values()) else (p[0] + "A", p[1]) for k, p in enumerate(pairs)]
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    res = gpu_lib.align_batch(scoring_to_params("lgotoh", sc), bases, off1, off2, len1, len2)
    for p, (a, b) in enumerate(pairs):
        if p in ub:
            assert int(res.ops_len[p]) == capi.PAIR_UNSUPPORTED
            continue
        o = orc.oracle_align("lgotoh", sc, a, b)
        assert int(res.score[p]) == o["score"] and np.array_equal(res.pair_ops(p), o["ops"]), p


def test_multi_device_uniform_batch_is_balanced(gpu_lib):
    """a uniform 1 M-pair-like batch over all visible devices gets ~equal cells per device."""
    nd = gpu_lib.device_count()
    if nd < 2:
        pytest.skip("one device visible")
    n = 400_000
    bases, off1, off2, l1, l2 = synth.batch_uniform(synth.SEED, 0, n, 150, 150)
    res = gpu_lib.align_batch(scoring_to_params("sw", S.linear(-1, 1, -1), device_count=nd, flags=capi.FLAG_OPS_2BIT), bases, off1, off2, l1, l2,
                              capi.Results(n, n * 76))
    split = gpu_lib.last_split().astype(np.float64)
    assert len(split) == nd and split.max() / split.mean() < 1.10, split
    one = gpu_lib.align_batch(scoring_to_params("sw", S.linear(-1, 1, -1), device_count=1, flags=capi.FLAG_OPS_2BIT), bases, off1, off2, l1, l2,
                              capi.Results(n, n * 76))
    assert np.array_equal(res.score, one.score) and np.array_equal(res.ops_len, one.ops_len)
    for p in (0, 1, n // 2, n - 1):
        assert np.array_equal(res.pair_ops(p), one.pair_ops(p))


@pytest.mark.parametrize("algo,sc", [("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("ggotoh", S.affine(-3, -1, 1, -1)),
                                     ("lgotoh", S.affine(-3, -1, 1, -1)), ("hirschberg", S.linear(-1, 2, -1)),
                                     ("myersmiller", S.affine(-3, -1, 1, -1))])
def test_two_bit_input_wire_format(gpu_lib, algo, sc):
    """SEQA_FLAG_BASES_2BIT (north_star: "sequences packed 2-bit/8-bit"): a ragged batch against the oracle, and dense /
    scattered packed layouts against the 8-bit form of the same call."""
    from test_emu_kernels import _same_results, shuffled_2bit_layout
    rng = np.random.default_rng(12)
    pairs = _defined(algo, random_pairs(rng, 600, 1, 300) + random_pairs(rng, 100, 1, 120, "AC") + random_pairs(rng, 6, 1500, 2500))
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, len1, len2)
    got = gpu_lib.align_batch(scoring_to_params(algo, sc, flags=capi.FLAG_BASES_2BIT), pk, p1, p2, len1, len2)
    assert compare_with_oracle_batch(got, algo, sc, bases, off1, off2, len1, len2, "2-bit in") == len(pairs)
    spk, q1, q2 = shuffled_2bit_layout(pk, p1, p2, len1, len2)
    got2 = gpu_lib.align_batch(scoring_to_params(algo, sc, flags=capi.FLAG_BASES_2BIT | capi.FLAG_OPS_2BIT), spk, q1, q2, len1, len2)
    want2 = gpu_lib.align_batch(scoring_to_params(algo, sc, flags=capi.FLAG_OPS_2BIT), bases, off1, off2, len1, len2)
    _same_results(want2, got2, len(pairs))


def test_two_bit_inputs_large_uniform_batch(gpu_lib):
    """300,000 x 150 bp through the one-shot call in several waves: 2-bit symbols in == 8-bit symbols in, both wire
    formats of the ops; the packed form moves 76 bytes per pair instead of 300."""
    n = 300_000
    sc = S.linear(-1, 1, -1)
    bases, off1, off2, l1, l2 = synth.batch_uniform(synth.SEED, 31_000_000, n, 150, 150)
    pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, l1, l2)
    assert len(pk) == n * 76
    want = gpu_lib.align_batch(scoring_to_params("sw", sc, flags=capi.FLAG_OPS_2BIT), bases, off1, off2, l1, l2, capi.Results(n, n * 76))
    got = gpu_lib.align_batch(scoring_to_params("sw", sc, flags=capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT), pk, p1, p2, l1, l2, capi.Results(n, n * 76))
    for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len", "ops_off"):
        assert np.array_equal(getattr(want, name), getattr(got, name)), name
    assert want.c.ops_used == got.c.ops_used and np.array_equal(want.ops[:want.c.ops_used], got.ops[:got.c.ops_used])
    rng = np.random.default_rng(3)
    for p in rng.integers(0, n, 300):
        p = int(p)
        a = bytes(bases[int(off1[p]):int(off1[p]) + 150]).decode()
        b = bytes(bases[int(off2[p]):int(off2[p]) + 150]).decode()
        o = orc.oracle_align("sw", sc, a, b)
        assert int(got.score[p]) == o["score"] and np.array_equal(got.pair_ops(p), o["ops"]), p


@pytest.mark.parametrize("algo,sc", [("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("ggotoh", S.affine(-3, -1, 1, -1)),
                                     ("lgotoh", S.affine(-3, -1, 1, -1)), ("hirschberg", S.linear(-1, 2, -1)),
                                     ("myersmiller", S.affine(-3, -1, 1, -1))])
def test_table_driven_symbol_equality(gpu_lib, algo, sc):
    """seqa_batch_in.sym_class: case-insensitive DNA with N matching nothing, purine / pyrimidine,
    protein groups -- identical to the oracle aligning the class-id strings with == (the reference's match(a, b) with the
    corresponding functor, include/SequenceAlignment.h:147).  A batch of mostly ACGT/acgt pairs stays on the packed
    kernels; only the pairs that hold another class run on the 8-bit kernels."""
    from test_emu_kernels import check_class_table, class_table_cases
    rng = np.random.default_rng(43)
    tables = dict(class_table_cases())
    for name, alphabet in (("dna-caseless-N-never", "ACGTacgtNnR"), ("purine-pyrimidine", "ACGTacgtN"), ("protein-groups", "AVLFWSTKRDEGPX"),
                           ("near-identity", "ACGT\xfe\xff#")):
        pairs = _defined(algo, random_pairs(rng, 300, 1, 250, alphabet) + random_pairs(rng, 60, 50, 300, alphabet, related=0.3))
        check_class_table(gpu_lib, algo, sc, pairs, tables[name])
    # mostly clean DNA in mixed case + a few pairs with N: the packed kernels take the batch, the N pairs are re-run one by one
    dna = tables["dna-caseless-N-never"]
    pairs = _defined(algo, random_pairs(rng, 400, 20, 200, "ACGTacgt"))
    for k in (3, 77, 200):
        pairs[k] = (pairs[k][0] + "N", "n" + pairs[k][1])
    check_class_table(gpu_lib, algo, sc, pairs, dna)
    if algo in ("sw", "nw", "ggotoh", "lgotoh"):
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        ctx = capi.Ctx(gpu_lib)
        ctx.upload(scoring_to_params(algo, sc), bases, off1, off2, len1, len2, sym_class=dna)
        ctx.run()
        ctx.sync()
        assert ctx.last_kernel().startswith("pk"), ctx.last_kernel()
        ctx.close()


def test_one_shot_call_speculative_first_wave(gpu_lib):
    """Batches above 524,288 pairs on one device start their first wave before the pass over the index arrays is complete
    (planned on pair 0's shape).  600,000 x 40 bp: (1) really uniform; (2) 50,000 pairs of another shape with the same slot
    size behind the first wave -- the helper thread refutes the assumption, the first wave stands, the rest goes through a
    second call; every pair of both against the threaded oracle, 8-bit and 2-bit symbols in."""
    n = 600_000
    sc = S.linear(-1, 1, -1)
    rng = np.random.default_rng(41)
    codes = rng.integers(0, 4, n * 80, dtype=np.uint8)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    for odd in (False, True):
        len1 = np.full(n, 40, np.uint32)
        len2 = np.full(n, 40, np.uint32)
        if odd:
            len1[400_000:450_000] = 33
            len2[400_000:450_000] = 47
        off1 = np.arange(n, dtype=np.uint64) * 80
        off2 = off1 + len1
        res = gpu_lib.align_batch(scoring_to_params("sw", sc), bases, off1, off2, len1, len2)
        compare_with_oracle_batch(res, "sw", sc, bases, off1, off2, len1, len2, label="speculative odd=%s" % odd)
        pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, len1, len2)
        got = gpu_lib.align_batch(scoring_to_params("sw", sc, flags=capi.FLAG_BASES_2BIT), pk, p1, p2, len1, len2)
        for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len", "ops_off"):
            assert np.array_equal(getattr(res, name), getattr(got, name)), (odd, name)
        assert np.array_equal(res.ops[:res.c.ops_used], got.ops[:got.c.ops_used])


@pytest.mark.gpu
@pytest.mark.parametrize("algo,sc", [("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("ggotoh", S.affine(-3, -1, 1, -1)),
                                     ("lgotoh", S.affine(-3, -1, 1, -1))])
def test_prep_tma_staging_and_global_loads(gpu_lib, algo, sc, monkeypatch):
    """pk_prep_kernel: a job's 128 sequences come in by one TMA bulk copy (cp.async.bulk + mbarrier) when they lie in one
    span of `bases` that fits the staging buffer, else by per-lane global loads -- dense, spread-out and mixed layouts of a
    uniform batch, with staging on and off, against the oracle."""
    rng = np.random.default_rng(92)
    for length in (150, 250):
        for bases, off1, off2, len1, len2, pairs in prep_staging_layouts(rng, 64 * 40 + 17, length):
            for tma in ("1", "0"):
                monkeypatch.setenv("SEQA_PREP_TMA", tma)
                res = gpu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
                compare_with_oracle_batch(res, algo, sc, bases, off1, off2, len1, len2, label="%s len %d tma %s" % (algo, length, tma))
