"""The oracle (oracle/seqa_oracle.c) pinned against the reference: README known answer, golden vectors produced
by the unmodified reference (oracle/make_golden.py) and, where the compiled reference is present, live
differential fuzzing."""
import numpy as np
import pytest

from common import orc, random_pairs

S = orc.Scoring


def sc_from(t):
    ctor, gap, go, ge, m, x, allow = t
    return S(ctor, gap=gap, gap_open=go, gap_extend=ge, match=m, mismatch=x, allow=bool(allow))


def test_readme_known_answer(oracle_built):
    # reference README.md:28-37 == test/Test.cpp:28-37
    o = orc.oracle_align("nw", S.linear(-1, 2), "AAAGAATGCAT", "AAACTCAT")
    r1, r2, fl = orc.expand("nw", "AAAGAATGCAT", "AAACTCAT", o["start_i"], o["start_j"], o["end_i"], o["end_j"], o["ops"])
    assert (r1, r2) == ("AAA-GAATGCAT", "AAAC---T-CAT")
    assert fl == "|||    | |||"
    o = orc.oracle_align("hirschberg", S.linear(-1, 2), "AAAGAATGCAT", "AAACTCAT")
    r1, r2, _ = orc.expand("hirschberg", "AAAGAATGCAT", "AAACTCAT", 0, 0, 11, 8, o["ops"])
    assert (r1, r2) == ("AAA-GAATGCAT", "AAAC---T-CAT")


def test_survey_known_answers(oracle_built):
    # SURVEY.md 8c golden vectors captured from the compiled reference
    P = ("AAAGAATGCAT", "AAACTCAT")
    cases = [("nw", S.linear(-1, 2, -1), "AAAGAATGCAT", "-AA-ACT-CAT"),
             ("hirschberg", S.linear(-1, 2, -1), "AAAGAATGCAT", "AAA-C-T-CAT"),
             ("sw", S.linear(-1, 2), "AAAGAA-TGCAT", "--A-AACT-CAT"),
             ("sw", S.linear(-2, 1, -1), "AAAGAATG-----CAT", "--------AAACTCAT"),
             ("ggotoh", S.affine(-3, -1, 1, -1), "AAAGAATGCAT", "A---AACTCAT"),
             ("ggotoh", S.affine(-3, -1, 1, -1, False), "AAA--GAATGCAT", "AAACT-----CAT"),
             ("lgotoh", S.affine(-3, -1, 2, -1), "AAAG-AATGCAT", "----AAACTCAT"),
             ("myersmiller", S.affine(-3, -1, 1, -1), "AAAGAATGCAT", "A---AACTCAT"),
             ("myersmiller", S.affine(-3, -1, 1, -1, False), "AAAGAA-TGCAT", "A---AACT-CAT")]
    for algo, sc, e1, e2 in cases:
        o = orc.oracle_align(algo, sc, *P)
        r1, r2, _ = orc.expand(algo, P[0], P[1], o["start_i"], o["start_j"], o["end_i"], o["end_j"], o["ops"])
        assert (r1, r2) == (e1, e2), (algo, sc)


def test_sw_score_access_vector(oracle_built):
    # std::mt19937 rng(1), interleaved draws, 150 bp, SW(-1,2,-1): MaxScore 109 at (143,149), 181 entries
    import random  # CPython's Mersenne Twister core == std::mt19937; getrandbits(32) gives the raw 32-bit outputs
    r = random.Random()
    r.setstate((3, tuple(_mt_init(1)) + (624,), None))
    a, b = [], []
    for _ in range(150):
        a.append("ACGT"[r.getrandbits(32) & 3])
        b.append("ACGT"[r.getrandbits(32) & 3])
    a, b = "".join(a), "".join(b)
    o = orc.oracle_align("sw", S.linear(-1, 2, -1), a, b)
    assert (o["score"], o["end_i"], o["end_j"]) == (109, 143, 149)
    r1, _, _ = orc.expand("sw", a, b, o["start_i"], o["start_j"], o["end_i"], o["end_j"], o["ops"])
    assert len(r1) == 181


def _mt_init(seed):
    mt = [0] * 624
    mt[0] = seed & 0xffffffff
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xffffffff
    return mt


def test_golden_vectors(oracle_built, golden):
    assert len(golden) >= 1500
    seen = set()
    for v in golden:
        sc = sc_from(v["scoring"])
        o = orc.oracle_align(v["algo"], sc, v["seq1"], v["seq2"])
        r1, r2, fl = orc.expand(v["algo"], v["seq1"], v["seq2"], o["start_i"], o["start_j"], o["end_i"], o["end_j"], o["ops"])
        assert (r1, r2, fl) == (v["row1"], v["row2"], v["flags"]), v
        if v["score"] is not None:
            assert o["score"] == v["score"], v
            if v["algo"] in ("sw", "lgotoh"):
                assert (o["end_i"], o["end_j"]) == (v["max_row"], v["max_col"]), v
        else:
            assert o["score"] == orc.rescore(v["algo"], sc, v["seq1"], v["seq2"], 0, 0, o["ops"]) or v["algo"] == "sw"
        seen.add(v["algo"])
    assert seen == set(orc.ALGOS)


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("algo", list(orc.ALGOS))
def test_differential_fuzz_vs_reference(oracle_built, algo):
    rng = np.random.default_rng(orc.ALGOS[algo] * 101 + 7)
    affine = algo in ("ggotoh", "lgotoh", "myersmiller")
    n = 0
    for rnd in range(12):
        allow = rng.random() > 0.25
        if affine:
            sc = S.affine(-int(rng.integers(0, 5)), -int(rng.integers(1, 4)), int(rng.integers(1, 5)), -int(rng.integers(1, 5)), allow)
        else:
            sc = S.linear(-int(rng.integers(1, 5)), int(rng.integers(1, 5)), -int(rng.integers(1, 5)), allow)
        alphabet = "ACGT" if rnd % 2 else "AC"
        for (a, b) in random_pairs(rng, 25, 1, 120, alphabet, related=0.4 if rnd % 3 == 0 else 0.0):
            if algo == "lgotoh" and (len(a), len(b)) in ((314, 288), (60, 57), (61, 58)):
                continue
            r = orc.ref_align(algo, sc, a, b)
            o = orc.oracle_align(algo, sc, a, b)
            got = orc.expand(algo, a, b, o["start_i"], o["start_j"], o["end_i"], o["end_j"], o["ops"])
            assert got == (r["row1"], r["row2"], r["flags"]), (algo, sc, a, b)
            if r["score"] is not None:
                assert o["score"] == r["score"]
            n += 1
    assert n > 250


def test_oracle_batch_entry_matches_single_pair_entry():
    """oracle_align_batch (threaded, used for the 100 k-pair GPU parity samples) == oracle_align pair by pair,
    and the vectorised comparison helper flags a single flipped op."""
    from common import _ragged_equal
    rng = np.random.default_rng(5)
    pairs = random_pairs(rng, 150, 0, 90) + random_pairs(rng, 50, 1, 200, related=0.3) + [("", "ACGT"), ("ACGT", ""), ("", "")]
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    for algo, sc in (("nw", orc.Scoring.linear(-1, 2, -1)), ("sw", orc.Scoring.linear(-1, 1, -1)),
                     ("ggotoh", orc.Scoring.affine(-3, -1, 1, -1)), ("lgotoh", orc.Scoring.affine(-3, -1, 1, -1, False)),
                     ("hirschberg", orc.Scoring.linear(-1, 2, -1)), ("myersmiller", orc.Scoring.affine(-3, -1, 1, -1))):
        r = orc.oracle_align_batch(algo, sc, bases, off1, off2, len1, len2, threads=3)
        for p, (a, b) in enumerate(pairs):
            o = orc.oracle_align(algo, sc, a, b)
            assert (int(r.score[p]), int(r.start_i[p]), int(r.start_j[p]), int(r.end_i[p]), int(r.end_j[p])) == \
                (o["score"], o["start_i"], o["start_j"], o["end_i"], o["end_j"]), (algo, p)
            assert np.array_equal(r.pair_ops(p), o["ops"]), (algo, p)
        assert _ragged_equal(r.ops, r.slot_off, r.ops, r.slot_off, r.ops_len) == -1
        flipped = r.ops.copy()
        victim = int(np.nonzero(r.ops_len > 3)[0][7])
        flipped[int(r.slot_off[victim]) + 2] ^= 3
        assert _ragged_equal(flipped, r.slot_off, r.ops, r.slot_off, r.ops_len) == victim
