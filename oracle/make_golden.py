"""TEST INFRASTRUCTURE.  Generates tests/golden/reference_vectors.json by running the UNMODIFIED reference
(oracle/_ref/libseqa_ref.so, built from /root/reference/include by oracle/Makefile) in the build container.
The GPU box has no /root/reference: tests there read only the committed JSON.

    python oracle/make_golden.py            # rewrites tests/golden/reference_vectors.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402

S = orc.Scoring
P = ("AAAGAATGCAT", "AAACTCAT")   # reference test/Test.cpp:28-29, README.md:28-37
Q = ("AATCG", "AACG")             # reference include/Test.cpp:36-37

LINEAR = [S.linear(-1, 2), S.linear(-1, 2, -1), S.linear(-2, 1, -1, False), S.linear(-1, 1, -1), S.linear(-2, 1, -1),
          S.linear(-3, 4, -2), S.linear(-4, 3, -4)]
AFFINE = [S.affine(-3, -1, 1, -1), S.affine(-3, -1, 2, -1), S.affine(-3, -1, 1, -1, False), S.affine(0, -1, 2, -3),
          S.affine(-4, -2, 3, -2), S.affine(-1, -3, 4, -1, False)]
ALGO_SCORINGS = {"nw": LINEAR, "sw": LINEAR, "hirschberg": LINEAR, "ggotoh": AFFINE, "lgotoh": AFFINE,
                 "myersmiller": AFFINE}
# reference include/Test.cpp:38-79 style hand pairs (written here, not copied) + tie-stress inputs
HAND = [P, Q, ("A", "A"), ("A", "C"), ("AAAA", "AAAA"), ("AAAAAAAA", "AAA"), ("ACACACAC", "CACACA"),
        ("GATTACA", "GCATGCT"), ("ACGTACGTACGT", "TGCATGCATGCA"), ("TTTTTTTTTTTTTTTT", "TTTTTTTT"),
        ("AAAAAAAAAA", "CCCCCCCCCC"), ("ACGT", "ACGT"), ("AGGTTGCCAT", "CAGGTTGACATT")]


def rand_pairs(rng, n, lo, hi, alphabet):
    out = []
    for _ in range(n):
        a = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), int(rng.integers(lo, hi + 1))))
        if rng.random() < 0.5:  # related pair: point edits of a
            b = []
            for ch in a:
                r = rng.random()
                if r < 0.15:
                    b.append(alphabet[int(rng.integers(0, len(alphabet)))])
                elif r < 0.22:
                    continue
                elif r < 0.29:
                    b += [ch, alphabet[int(rng.integers(0, len(alphabet)))]]
                else:
                    b.append(ch)
            b = "".join(b) or alphabet[0]
        else:
            b = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), int(rng.integers(lo, hi + 1))))
        out.append((a, b))
    return out


def undefined_in_reference(algo, s1, s2):
    if algo == "lgotoh":
        if len(s1) == 0 or len(s2) == 0:
            return True  # reads uninitialised MaxRow/MaxCol (reference include/SALocalGotoh.h:30-31,285)
        if (len(s1), len(s2)) in ((314, 288), (60, 57), (61, 58)):
            return True  # reference include/SALocalGotoh.h:484-488
    return False


def main():
    if not orc.have_ref():
        raise SystemExit("oracle/_ref/libseqa_ref.so missing: run `make -C oracle` in the build container")
    rng = np.random.default_rng(20240607)
    vectors = []
    for algo, scorings in ALGO_SCORINGS.items():
        pairs = list(HAND) + [("", P[1]), (P[0], ""), ("", "")]
        pairs += rand_pairs(rng, 14, 1, 40, "ACGT") + rand_pairs(rng, 8, 1, 30, "AC") + rand_pairs(rng, 4, 100, 170, "ACGT")
        for sc in scorings:
            for (a, b) in pairs:
                if undefined_in_reference(algo, a, b):
                    continue
                r = orc.ref_align(algo, sc, a, b, functor=False)
                vectors.append(dict(algo=algo, scoring=list(sc.astuple()), seq1=a, seq2=b, row1=r["row1"],
                                    row2=r["row2"], flags=r["flags"], score=r["score"],
                                    max_row=r["max_row"], max_col=r["max_col"]))
    # the survey's score-access check: mt19937(1) interleaved draws, 150 bp, SW (-1,2,-1)
    out = dict(generator="oracle/make_golden.py", reference="przemektmalon/SeqALib include/*.h (unmodified, compiled by path)",
               n=len(vectors), vectors=vectors)
    dst = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d vectors, %d bytes" % (dst, len(vectors), os.path.getsize(dst)))


if __name__ == "__main__":
    main()
