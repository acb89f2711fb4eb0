"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/seqa_cuda.h declares, and (on a
box without a GPU) refuses to compute instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, capi, orc, scoring_to_params


def _built():
    if not os.path.exists(capi.DEFAULT_SO):
        import __graft_entry__ as g
        g.build()
    return capi.Lib()


def test_exports_match_header():
    lib = _built()
    hdr = open(os.path.join(ROOT, "include", "seqa_cuda.h")).read()
    declared = set(re.findall(r"\b(seqa_(?:cuda|ctx)_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib.L, name), name
    assert lib.L.seqa_cuda_abi_version() == 2


def test_no_torch_types_in_signatures():
    hdr = open(os.path.join(ROOT, "include", "seqa_cuda.h")).read()
    assert "torch" not in hdr.lower() and "at::" not in hdr and "#include <stdint.h>" in hdr


def test_no_cpu_fallback_without_device():
    lib = _built()
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    bases, off1, off2, len1, len2 = orc.batch_arrays([("ACGT", "ACGA")])
    prm = scoring_to_params("nw", orc.Scoring.linear(-1, 2))
    with pytest.raises(capi.SeqaError) as e:
        lib.align_batch(prm, bases, off1, off2, len1, len2)
    assert e.value.code == -3  # SEQA_ERR_NO_DEVICE
    h = C.c_void_p()
    assert lib.L.seqa_ctx_create(C.byref(h), 0, None) == -3


def test_product_does_not_reference_oracle():
    # the product path (package + headers) must never import / link / execute anything under oracle/
    for base in ("seqalib_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".hpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "pyoracle" not in txt and "seqa_oracle" not in txt and "libseqa_ref" not in txt, f


def test_param_validation(emu_lib):
    bases, off1, off2, len1, len2 = orc.batch_arrays([("ACGT", "ACGA")])
    for bad in (capi.make_params("nw", gap=1, match=1), capi.make_params("nw", gap=-1, match=0),
                capi.make_params("ggotoh", gap_open=1, gap_extend=-1, match=1),
                capi.make_params("sw", gap=-1, match=1, mismatch=2)):
        with pytest.raises(capi.SeqaError) as e:
            emu_lib.align_batch(bad, bases, off1, off2, len1, len2)
        assert e.value.code == -2
    # offsets that would wrap around 2^64 are caught, not added
    o_bad = off1.copy()
    o_bad[0] = np.uint64(2 ** 64 - 2)
    with pytest.raises(capi.SeqaError) as e:
        emu_lib.align_batch(scoring_to_params("nw", orc.Scoring.linear(-1, 2)), bases, o_bad, off2, len1, len2)
    assert e.value.code == -1
    small = capi.Results(1, 2)
    with pytest.raises(capi.SeqaError) as e:
        emu_lib.align_batch(scoring_to_params("nw", orc.Scoring.linear(-1, 2)), bases, off1, off2, len1, len2, small)
    assert e.value.code == -5


def test_local_gotoh_undefined_shapes_are_rejected_per_pair(emu_lib):
    """The three LocalGotoh shapes that are undefined behaviour in the reference (include/SALocalGotoh.h:484-488) come
    back with ops_len = SEQA_PAIR_UNSUPPORTED; every other pair of the batch is aligned and matches the oracle."""
    rng = np.random.default_rng(8)
    sc = orc.Scoring.affine(-3, -1, 1, -1)

    def seq(n):
        return "".join("ACGT"[k] for k in rng.integers(0, 4, n))
    pairs = [(seq(40), seq(45)), (seq(60), seq(57)), (seq(33), seq(20)), (seq(61), seq(58)), (seq(57), seq(60)), (seq(314), seq(288)),
             (seq(70), seq(9))]
    ub = {1, 3, 5}
    for flags in (0, capi.FLAG_FORCE_GENERIC, capi.FLAG_OPS_2BIT):
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        res = emu_lib.align_batch(scoring_to_params("lgotoh", sc, flags=flags), bases, off1, off2, len1, len2)
        for p, (a, b) in enumerate(pairs):
            if p in ub:
                assert int(res.ops_len[p]) == capi.PAIR_UNSUPPORTED and int(res.score[p]) == -2 ** 31
                continue
            o = orc.oracle_align("lgotoh", sc, a, b)
            assert int(res.score[p]) == o["score"] and np.array_equal(res.pair_ops(p), o["ops"]), (flags, p)
    # a uniform batch of such a shape: every pair rejected, the call still succeeds
    uni = [(seq(60), seq(57)) for _ in range(5)]
    res = emu_lib.align_batch(scoring_to_params("lgotoh", sc), *orc.batch_arrays(uni))
    assert (res.ops_len[:5] == capi.PAIR_UNSUPPORTED).all() and res.c.ops_used == 0
    # the same shapes are ordinary for every other algorithm
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    g = emu_lib.align_batch(scoring_to_params("ggotoh", sc), bases, off1, off2, len1, len2)
    assert (g.ops_len[:len(pairs)] != capi.PAIR_UNSUPPORTED).all()


def test_download_range_matches_full_download(emu_lib):
    rng = np.random.default_rng(4)
    from common import random_pairs
    pairs = random_pairs(rng, 37, 1, 60)
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    for flags in (0, capi.FLAG_OPS_2BIT):
        ctx = capi.Ctx(emu_lib)
        ctx.upload(scoring_to_params("sw", orc.Scoring.linear(-1, 1, -1), flags=flags), bases, off1, off2, len1, len2)
        ctx.run()
        full = ctx.download()
        for first, count in ((0, 37), (5, 9), (36, 1), (11, 0)):
            part = ctx.download_range(first, count, 4096)
            for k in range(count):
                for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
                    assert getattr(part, name)[k] == getattr(full, name)[first + k], (name, first, k)
                assert np.array_equal(part.pair_ops(k), full.pair_ops(first + k))
        with pytest.raises(capi.SeqaError):
            ctx.download_range(30, 8, 4096)
        ctx.close()


def test_multi_device_split_is_balanced(emu_lib):
    """a uniform batch over several devices is cut so that every device gets (nearly) the same cells --
    the whole-round wave sizing must not override the at-least-two-waves-per-device cap."""
    n = 4000
    pairs = [("ACGTACGTAC", "ACGTTCGTAC")] * n
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    res = emu_lib.align_batch(scoring_to_params("sw", orc.Scoring.linear(-1, 1, -1), device_count=2), bases, off1, off2, len1, len2)
    split = emu_lib.last_split().astype(np.float64)
    assert len(split) == 2 and split.sum() == n * 101
    assert split.max() / split.mean() < 1.10, split
    o = orc.oracle_align("sw", orc.Scoring.linear(-1, 1, -1), *pairs[0])
    assert (res.score[:n] == o["score"]).all()


def test_wave_schedule_of_uniform_batches(emu_lib, monkeypatch):
    """One-shot call on a uniform batch larger than two rounds of the fill kernel (emulator: 1,536 pairs per round): a
    one-round first wave, multi-round waves behind it, ops of every wave at its own base offset -- results must not
    depend on the schedule (SEQA_WAVE_ROUNDS 1 / default, both wire formats, explicit SEQA_WAVE_MCELLS)."""
    rng = np.random.default_rng(77)
    n, L = 5000, 10
    codes = rng.integers(0, 4, (n, 2 * L))
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1).copy()
    off1 = np.arange(n, dtype=np.uint64) * np.uint64(2 * L)
    off2 = off1 + np.uint64(L)
    l1 = np.full(n, L, np.uint32)
    l2 = np.full(n, L, np.uint32)
    sc = orc.Scoring.linear(-1, 1, -1)
    want = orc.oracle_align_batch("sw", sc, bases, off1, off2, l1, l2)
    pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, l1, l2)
    for env in ({}, {"SEQA_WAVE_ROUNDS": "1"}, {"SEQA_WAVE_MCELLS": "10"}, {"SEQA_TWO_COMPUTE_STREAMS": "1"}):
        for k in ("SEQA_WAVE_ROUNDS", "SEQA_WAVE_MCELLS", "SEQA_TWO_COMPUTE_STREAMS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for flags, ins in ((0, (bases, off1, off2, l1, l2)), (capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT, (pk, p1, p2, l1, l2))):
            got = emu_lib.align_batch(scoring_to_params("sw", sc, flags=flags), *ins)
            for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
                assert np.array_equal(getattr(got, name)[:n], getattr(want, name)[:n]), (env, flags, name)
            for p in list(range(0, n, 97)) + [1535, 1536, 1537, n - 1]:
                assert np.array_equal(got.pair_ops(p), want.pair_ops(p)), (env, flags, p)
