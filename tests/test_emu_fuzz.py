"""Differential fuzz of the product's kernel sources (built for the CPU SIMT emulator, tests/emu) against the C oracle:
random scorings inside the packed domain and outside it, uniform / ragged / related batches, every aligner, 8-bit and
2-bit symbols in, byte and 2-bit ops out, one and two (emulated) devices, packed and forced-generic kernels.  Bounded by
wall time so the CPU tier stays a few minutes; `python tests/test_emu_fuzz.py SEED SECONDS` runs it for longer
(at the end of round 2: 10,700 batches of the matrix aligners, 950 with 2-bit symbols / two devices / forced-generic kernels /
linear-space aligners in ad-hoc forms of this script, and 1,150 batches of this exact form, all without a mismatch)."""
import sys
import time

import numpy as np

from common import EMU_SO, capi, check_batch_against_oracle, orc, random_pairs, scoring_to_params

S = orc.Scoring
UB_SHAPES = ((314, 288), (60, 57), (61, 58))  # LocalGotoh shapes the reference overrides (include/SALocalGotoh.h:484-488)


def _same(want, got, n):
    for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
        assert np.array_equal(getattr(want, name)[:n], getattr(got, name)[:n]), name
    for p in range(n):
        assert np.array_equal(want.pair_ops(p), got.pair_ops(p)), p


def fuzz(lib, seed, seconds, linear_space=True, long_len=(250, 400)):
    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds
    n = 0
    algos = ["nw", "sw", "ggotoh", "lgotoh"] + (["hirschberg", "myersmiller"] if linear_space else [])
    while time.time() < t_end:
        n += 1
        algo = algos[rng.integers(0, len(algos))]
        allow = bool(rng.integers(0, 4) != 0)
        if algo in ("nw", "sw", "hirschberg"):
            sc = S.linear(-int(rng.integers(1, 5)), int(rng.integers(1, 5)), -int(rng.integers(1, 5)), allow)
        else:
            sc = S.affine(-int(rng.integers(0, 5)), -int(rng.integers(1, 4)), int(rng.integers(1, 5)), -int(rng.integers(1, 5)), allow)
        ls = algo in ("hirschberg", "myersmiller")
        kind = int(rng.integers(0, 3))
        if kind == 0:  # uniform: the TMA-staged prep, identity permutation, whole-round planning
            L = int(rng.integers(1, 60))
            pairs = random_pairs(rng, int(rng.integers(8, 20)) if ls else int(rng.integers(64, 200)), L, L)
        elif kind == 1:  # ragged, with empty sequences
            pairs = random_pairs(rng, int(rng.integers(6, 16)) if ls else int(rng.integers(40, 150)), 0, 70, related=float(rng.choice([0.0, 0.3])))
        else:  # a few longer pairs among short ones
            pairs = random_pairs(rng, 2 if ls else int(rng.integers(3, 10)), long_len[0], long_len[1], related=0.2) + random_pairs(rng, 8 if ls else 30, 1, 40)
        if algo == "lgotoh":
            pairs = [p for p in pairs if p[0] and p[1] and (len(p[0]), len(p[1])) not in UB_SHAPES]
        flags = int(rng.choice([0, capi.FLAG_OPS_2BIT])) | int(rng.choice([0, 0, 0, capi.FLAG_FORCE_GENERIC]))
        dc = int(rng.choice([1, 1, 2]))
        label = "fuzz seed %d batch %d" % (seed, n)
        check_batch_against_oracle(lib, algo, sc, pairs, flags=flags, device_count=dc, label=label)
        if rng.integers(0, 3) == 0:  # the same batch with 2-bit symbols in
            bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
            want = lib.align_batch(scoring_to_params(algo, sc, flags=flags, device_count=dc), bases, off1, off2, len1, len2)
            pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, len1, len2)
            got = lib.align_batch(scoring_to_params(algo, sc, flags=flags | capi.FLAG_BASES_2BIT, device_count=dc), pk, p1, p2, len1, len2)
            _same(want, got, len(pairs))
    return n


def test_emulated_kernels_against_oracle_fuzz(emu_lib):
    assert fuzz(emu_lib, 20261019, 30.0, long_len=(90, 140)) >= 5  # a mismatch raises inside; the count only shows it ran


if __name__ == "__main__":
    orc.build()
    print("batches:", fuzz(capi.Lib(EMU_SO), int(sys.argv[1]) if len(sys.argv) > 1 else 1, float(sys.argv[2]) if len(sys.argv) > 2 else 300.0))
