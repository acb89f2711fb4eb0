"""Kernel LOGIC on the CPU: the unmodified kernel sources, compiled for the SIMT emulator under tests/emu, run
through the same C ABI and are compared with the oracle.  (The real parity tests are the `-m gpu` ones; this keeps
index arithmetic, tie-breaks and the packed-integer tricks checked where no GPU exists.)"""
import numpy as np
import pytest

from common import capi, check_batch_against_oracle, orc, prep_staging_layouts, random_pairs, scoring_to_params
from test_oracle_golden import sc_from

S = orc.Scoring
EDGE = [("AAAGAATGCAT", "AAACTCAT"), ("AATCG", "AACG"), ("", "ACGT"), ("ACGT", ""), ("", ""), ("A", "A"), ("A", "C"),
        ("AAAAAAAAAAAAAAAAAAAA", "AAAAA"), ("ACGTNACGT", "ACGTACGT"), ("acgt", "ACGT")]


@pytest.mark.parametrize("algo,sc", [
    ("nw", S.linear(-1, 2)), ("nw", S.linear(-1, 2, -1)), ("nw", S.linear(-3, 4, -2)),
    ("sw", S.linear(-1, 1, -1)), ("sw", S.linear(-2, 1, -1, False)), ("sw", S.linear(-1, 2)),
    ("ggotoh", S.affine(-3, -1, 1, -1)), ("ggotoh", S.affine(-3, -1, 1, -1, False)), ("ggotoh", S.affine(0, -2, 3, -1)),
    ("lgotoh", S.affine(-3, -1, 1, -1)), ("lgotoh", S.affine(-1, -1, 2, -2, False)),
])
def test_matrix_algorithms(emu_lib, algo, sc):
    rng = np.random.default_rng(11)
    pairs = list(EDGE) + random_pairs(rng, 30, 1, 70) + random_pairs(rng, 8, 1, 50, "AC") + \
        random_pairs(rng, 6, 120, 180) + random_pairs(rng, 8, 1, 90, related=0.3)
    for flags in (0, capi.FLAG_TRACE8, capi.FLAG_FORCE_GENERIC, capi.FLAG_OPS_2BIT):  # OPS_2BIT: 4 ops per byte on the wire
        check_batch_against_oracle(emu_lib, algo, sc, pairs, flags=flags)


@pytest.mark.parametrize("algo,sc", [
    ("hirschberg", S.linear(-1, 2, -1)), ("hirschberg", S.linear(-2, 3, -2, False)),
    ("myersmiller", S.affine(-3, -1, 1, -1)), ("myersmiller", S.affine(-3, -1, 1, -1, False)),
])
def test_linear_space_algorithms(emu_lib, algo, sc):
    """expand / sweep / split kernels incl. the row-block pipeline between warps (FLAG_LS_R1 = 32-row blocks, so a
    150-row sweep is 5 tasks deep; the emulator's fibers yield inside the progress-counter spin)."""
    rng = np.random.default_rng(12)
    pairs = list(EDGE) + random_pairs(rng, 10, 1, 70) + random_pairs(rng, 4, 1, 50, "AC") + \
        random_pairs(rng, 2, 150, 300) + random_pairs(rng, 3, 1, 200, related=0.3)
    check_batch_against_oracle(emu_lib, algo, sc, pairs)  # contains non-ACGT symbols: int32 sweeps
    check_batch_against_oracle(emu_lib, algo, sc, pairs[-12:], flags=capi.FLAG_LS_R1)
    clean = [p for p in pairs if set(p[0] + p[1]) <= set("ACGT")]  # packed forward+reverse s16x2 sweeps
    check_batch_against_oracle(emu_lib, algo, sc, clean)
    check_batch_against_oracle(emu_lib, algo, sc, clean[-12:], flags=capi.FLAG_LS_R1 | capi.FLAG_OPS_2BIT)


def test_packed_path_is_taken(emu_lib):
    rng = np.random.default_rng(3)
    pairs = random_pairs(rng, 70, 20, 40)
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    ctx = capi.Ctx(emu_lib)
    ctx.upload(scoring_to_params("sw", S.linear(-1, 1, -1)), bases, off1, off2, len1, len2)
    ctx.run()
    ctx.sync()
    assert ctx.last_kernel() == "pk_fill_sw_s16x2_t2"  # default SW scoring: 2 trace bits per cell
    assert ctx.cells() == int((len1.astype(np.int64) * len2).sum())
    # a non-ACGT base voids the packed result of ITS pair only: that pair is re-run on the 8-bit kernels, the others keep
    # their packed results; later runs of the resident batch plan it onto the 8-bit kernels from the start
    pairs[5] = (pairs[5][0][:10] + "N" + pairs[5][0][10:], pairs[5][1])
    pairs[66] = (pairs[66][0], "n" + pairs[66][1])
    for algo, sc in (("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("ggotoh", S.affine(-3, -1, 1, -1))):
        check_batch_against_oracle(emu_lib, algo, sc, pairs)
        check_batch_against_oracle(emu_lib, algo, sc, pairs, flags=capi.FLAG_OPS_2BIT, device_count=2)
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        ctx = capi.Ctx(emu_lib)
        ctx.upload(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
        ctx.run()
        first = ctx.download()
        assert ctx.last_kernel().startswith("pk")  # the batch stayed on the packed kernels
        ctx.run()
        second = ctx.download()
        ctx.close()
        for p, (a, b) in enumerate(pairs):
            o = orc.oracle_align(algo, sc, a, b)
            for res in (first, second):
                assert int(res.score[p]) == o["score"] and np.array_equal(res.pair_ops(p), o["ops"]), (algo, p)


def test_golden_through_emulated_kernels(emu_lib, golden):
    groups = {}
    for v in golden:
        if v["algo"] in ("hirschberg", "myersmiller") and False:
            continue
        groups.setdefault((v["algo"], tuple(v["scoring"])), []).append(v)
    n = 0
    for (algo, sct), vs in groups.items():
        sc = sc_from(sct)
        pairs = [(v["seq1"], v["seq2"]) for v in vs]
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        res = emu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
        for p, v in enumerate(vs):
            got = orc.expand(algo, v["seq1"], v["seq2"], int(res.start_i[p]), int(res.start_j[p]), int(res.end_i[p]),
                             int(res.end_j[p]), res.pair_ops(p))
            assert got == (v["row1"], v["row2"], v["flags"]), v
            if v["score"] is not None:
                assert int(res.score[p]) == v["score"], v
            n += 1
    assert n == len(golden)


def test_two_device_shards(emu_lib):
    # the emulator exposes two fake devices: exercises the static split + gather of seqa_cuda_align_batch
    rng = np.random.default_rng(5)
    pairs = random_pairs(rng, 41, 1, 60)
    check_batch_against_oracle(emu_lib, "nw", S.linear(-1, 2, -1), pairs, device_count=2)
    check_batch_against_oracle(emu_lib, "lgotoh", S.affine(-3, -1, 1, -1), pairs, device_count=2)
    check_batch_against_oracle(emu_lib, "sw", S.linear(-1, 1, -1), pairs, flags=capi.FLAG_OPS_2BIT, device_count=2)


def test_synthetic_generator_matches_spec(emu_lib):
    # SURVEY.md 8d generator: regenerate on the CPU, compare with what the device wrote
    def splitmix64(z):
        z = (z + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        return z ^ (z >> 31)
    seed, first, n = 20240607, 1000, 9
    ctx = capi.Ctx(emu_lib)
    ctx.generate(scoring_to_params("sw", S.linear(-1, 1, -1)), seed, first, n, 1)
    l1 = [50 + splitmix64(splitmix64(seed ^ (2 * (first + p))) ^ 0xC0FFEE) % 951 for p in range(n)]
    l2 = [50 + splitmix64(splitmix64(seed ^ (2 * (first + p) + 1)) ^ 0xC0FFEE) % 951 for p in range(n)]
    bases, off1, off2, len1, len2 = ctx.download_inputs(sum(l1) + sum(l2))
    assert len1.tolist() == l1 and len2.tolist() == l2
    for p in range(n):
        for w, (off, ln) in enumerate(((off1[p], l1[p]), (off2[p], l2[p]))):
            key = splitmix64(seed ^ (2 * (first + p) + w))
            exp = "".join("ACGT"[(splitmix64((key + pos // 32) & (2 ** 64 - 1)) >> (2 * (pos % 32))) & 3] for pos in range(ln))
            assert bytes(bases[int(off):int(off) + ln]).decode() == exp


@pytest.mark.parametrize("variant", [1, 2])
def test_measured_trace_layout_variants_stay_correct(oracle_built, tmp_path, variant):
    """PK_PAIR_PIECES=1/2 (the sector-private trace layouts measured and rejected in profiles/r01_notes_walk_layout.md)
    are kept behind a compile-time switch: build each for the emulator and check the packed NW/SW path still matches."""
    import os
    import subprocess
    from common import ROOT
    so = str(tmp_path / ("libseqa_emu_pp%d.so" % variant))
    csrc = os.path.join(ROOT, "seqalib_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-DSEQA_EMU", "-DPK_PAIR_PIECES=%d" % variant,
                           "-I" + os.path.join(ROOT, "tests", "emu"), "-I" + csrc, "-Wno-unknown-pragmas", "-x", "c++",
                           os.path.join(csrc, "seqa_cuda.cu"), os.path.join(ROOT, "tests", "emu", "cuda_emu.cpp"), "-o", so])
    lib = capi.Lib(so)
    rng = np.random.default_rng(21)
    pairs = random_pairs(rng, 40, 1, 70) + random_pairs(rng, 5, 120, 180) + random_pairs(rng, 8, 1, 90, related=0.3) + \
        [("", ""), ("A", "C"), ("ACGT", "")]
    for algo, sc in (("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-2, 1, -1, False))):
        check_batch_against_oracle(lib, algo, sc, pairs)


@pytest.mark.parametrize("algo,sc", [("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("ggotoh", S.affine(-3, -1, 1, -1))])
def test_prep_staging_and_global_loads_agree(emu_lib, algo, sc, monkeypatch):
    """pk_prep_kernel stages a job's span of `bases` in shared memory by one TMA bulk copy when it fits and reads global
    memory otherwise; SEQA_PREP_TMA=0 never stages.  All of them give the oracle's alignments."""
    rng = np.random.default_rng(91)
    for bases, off1, off2, len1, len2, pairs in prep_staging_layouts(rng, 64 * 2 + 9, 33):
        want = orc.oracle_align_batch(algo, sc, bases, off1, off2, len1, len2)
        for tma in ("1", "0"):
            monkeypatch.setenv("SEQA_PREP_TMA", tma)
            got = emu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
            _same_results(want, got, len(pairs))


@pytest.mark.parametrize("flags", [0, capi.FLAG_OPS_2BIT])
def test_gather_ops_word_path(emu_lib, flags):
    """gather_ops_kernel with more pairs than warps (8 lanes per pair): op strings of every length and alignment, copied as
    words of the destination grid (byte ops) or packed 16 ops per word (2-bit ops), first / last word byte by byte."""
    rng = np.random.default_rng(55)
    pairs = random_pairs(rng, 1500, 0, 45) + random_pairs(rng, 40, 60, 120, related=0.2)
    for algo, sc in (("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1))):
        check_batch_against_oracle(emu_lib, algo, sc, pairs, flags=flags)


def test_uniform_batch_cut_at_whole_rounds(emu_lib, monkeypatch):
    """A uniform batch that outgrows the scratch budget is cut into chunks of whole fill rounds (one job per resident warp:
    2 SMs x 3 CTAs x 4 warps = 24 jobs on the emulated device); results do not depend on where the cuts fall."""
    rng = np.random.default_rng(77)
    pairs = random_pairs(rng, 64 * 58 + 5, 18, 18)  # 59 jobs of 64 pairs
    # budgets of ~30 jobs of this shape per chunk -> cut at 24 (GlobalGotoh: 3 x 4-bit planes; SW: 2-bit trace; NW: 4-bit)
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    for algo, sc, kb in (("ggotoh", S.affine(-3, -1, 1, -1), "2200"), ("sw", S.linear(-1, 1, -1), "850"), ("nw", S.linear(-1, 2, -1), "1150")):
        monkeypatch.setenv("SEQA_SCRATCH_BUDGET_KB", kb)
        want = orc.oracle_align_batch(algo, sc, bases, off1, off2, len1, len2)
        for rounds in ("1", "0"):
            monkeypatch.setenv("SEQA_ROUND_CHUNKS", rounds)
            got = emu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
            _same_results(want, got, len(pairs))


def test_batch_layouts_dense_and_scattered(emu_lib):
    """A dense batch (seq1, seq2, next pair ... back to back) sends no offset arrays -- the device derives them from the
    op-slot scan -- and a uniform one no length arrays either; any other layout (gaps, reordered or shared sequences,
    a dense run that does not start at byte 0) must give the same alignments."""
    rng = np.random.default_rng(33)
    sc = S.linear(-1, 1, -1)
    for pairs in (random_pairs(rng, 80, 1, 90), [("ACGTACGTAC" * 3, "ACGTTCGTAC" * 3)] * 70):  # ragged / uniform
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        want = emu_lib.align_batch(scoring_to_params("sw", sc), bases, off1, off2, len1, len2)
        # (1) dense, but behind 5 bytes of padding; (2) pairs stored in reverse order with a gap byte between sequences
        pad = np.concatenate([np.frombuffer(b"NNNNN", dtype=np.uint8), bases])
        layouts = [(pad, off1 + np.uint64(5), off2 + np.uint64(5))]
        chunks, o1, o2, pos = [], np.zeros(len(pairs), np.uint64), np.zeros(len(pairs), np.uint64), 0
        for p in reversed(range(len(pairs))):
            a, b = pairs[p]
            o2[p], o1[p] = pos, pos + len(b) + 1
            chunks += [b.encode(), b"N", a.encode(), b"N"]
            pos += len(a) + len(b) + 2
        layouts.append((np.frombuffer(b"".join(chunks), dtype=np.uint8), o1, o2))
        for lb, lo1, lo2 in layouts:
            got = emu_lib.align_batch(scoring_to_params("sw", sc), np.ascontiguousarray(lb), lo1, lo2, len1, len2)
            for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
                assert np.array_equal(getattr(want, name), getattr(got, name)), name
            for p in range(len(pairs)):
                assert np.array_equal(want.pair_ops(p), got.pair_ops(p)), p
        check_batch_against_oracle(emu_lib, "sw", sc, pairs)


def _same_results(want, got, n):
    for name in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
        assert np.array_equal(getattr(want, name)[:n], getattr(got, name)[:n]), name
    for p in range(n):
        assert np.array_equal(want.pair_ops(p), got.pair_ops(p)), p


def shuffled_2bit_layout(pk, p1, p2, len1, len2):
    """the same packed sequences in a NON-dense layout: junk in front, pairs in reverse order, sequence 2 before
    sequence 1, a junk byte between them (exercises the offset-array path of SEQA_FLAG_BASES_2BIT)"""
    n = len(len1)
    chunks, q1, q2, pos = [b"\xff\xff\xff"], np.zeros(n, np.uint64), np.zeros(n, np.uint64), 3
    for p in reversed(range(n)):
        b1, b2 = (int(len1[p]) + 3) // 4, (int(len2[p]) + 3) // 4
        q2[p] = pos
        chunks += [pk[int(p2[p]):int(p2[p]) + b2].tobytes(), b"\xaa"]
        pos += b2 + 1
        q1[p] = pos
        chunks.append(pk[int(p1[p]):int(p1[p]) + b1].tobytes())
        pos += b1
    return np.frombuffer(b"".join(chunks), dtype=np.uint8).copy(), q1, q2


@pytest.mark.parametrize("algo,sc", [("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("ggotoh", S.affine(-3, -1, 1, -1)),
                                     ("lgotoh", S.affine(-3, -1, 1, -1)), ("hirschberg", S.linear(-1, 2, -1)),
                                     ("myersmiller", S.affine(-3, -1, 1, -1))])
def test_two_bit_input_wire_format(emu_lib, algo, sc):
    """SEQA_FLAG_BASES_2BIT: 2-bit symbols in (dense ragged, dense uniform and scattered layouts; several devices) give
    exactly the results of the 8-bit form, which is checked against the oracle."""
    rng = np.random.default_rng(5)
    ragged = random_pairs(rng, 16, 1, 50) + [("", "ACGT"), ("ACGT", ""), ("", ""), ("A", "C")] + random_pairs(rng, 2, 90, 130)
    if algo == "lgotoh":
        ragged = [p for p in ragged if p[0] and p[1]]
    uniform = [("ACGTACGTACGTTGCAAC", "ACGTTCGTACGGGTTGCAATCA")] * 66
    if algo in ("hirschberg", "myersmiller"):  # the emulated recursion is slow: a handful of pairs shows the format works
        ragged, uniform = ragged[:4] + ragged[16:22], uniform[:6]
    check_batch_against_oracle(emu_lib, algo, sc, ragged)
    for pairs in (ragged, uniform):
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, len1, len2)
        for flags, dc in ((0, 1), (capi.FLAG_OPS_2BIT | capi.FLAG_FORCE_GENERIC, 2)):
            want = emu_lib.align_batch(scoring_to_params(algo, sc, flags=flags, device_count=dc), bases, off1, off2, len1, len2)
            got = emu_lib.align_batch(scoring_to_params(algo, sc, flags=flags | capi.FLAG_BASES_2BIT, device_count=dc), pk, p1, p2, len1, len2)
            _same_results(want, got, len(pairs))
        spk, q1, q2 = shuffled_2bit_layout(pk, p1, p2, len1, len2)
        want = emu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
        got = emu_lib.align_batch(scoring_to_params(algo, sc, flags=capi.FLAG_BASES_2BIT), spk, q1, q2, len1, len2)
        _same_results(want, got, len(pairs))
    with pytest.raises(ValueError):
        capi.pack_bases_2bit(*orc.batch_arrays([("ACGNT", "ACGT")]))
    # a packed buffer that is too short for its offsets is refused
    bases, off1, off2, len1, len2 = orc.batch_arrays(ragged)
    pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, len1, len2)
    with pytest.raises(capi.SeqaError):
        emu_lib.align_batch(scoring_to_params(algo, sc, flags=capi.FLAG_BASES_2BIT), pk[:len(pk) // 2].copy(), p1, p2, len1, len2)


def class_table_cases():
    """(name, sym_class uint8[256]) -- case-insensitive DNA with 'N' / 'n' matching nothing; purine / pyrimidine; a
    20-letter protein alphabet folded into 6 groups (everything else: matches nothing)."""
    ident = np.arange(256, dtype=np.uint8)
    dna = np.full(256, 255, dtype=np.uint8)
    for k, (u, l) in enumerate(zip(b"ACGT", b"acgt")):
        dna[u] = dna[l] = k
    for k, ch in enumerate(b"RYKMSW"):  # IUPAC codes: each its own class here
        dna[ch] = 4 + k
    ry = np.full(256, 255, dtype=np.uint8)
    for ch in b"AGag":
        ry[ch] = 0
    for ch in b"CTct":
        ry[ch] = 1
    prot = np.full(256, 255, dtype=np.uint8)
    for k, grp in enumerate((b"AVLIMC", b"FWYH", b"STNQ", b"KR", b"DE", b"GP")):
        for ch in grp:
            prot[ch] = k
    ident2 = ident.copy()
    ident2[254] = ident2[255] = 253  # ids must stay below 254: fold the two top bytes into one class
    return [("dna-caseless-N-never", dna), ("purine-pyrimidine", ry), ("protein-groups", prot), ("near-identity", ident2)]


def translate_for_oracle(table, s, which):
    """the string the oracle aligns with == to reproduce the class table: class ids as bytes, the 'matches nothing'
    class as two different bytes in the two sequences"""
    return bytes((254 + which) if table[c] == 255 else int(table[c]) for c in s.encode("latin1")).decode("latin1")


def check_class_table(lib, algo, sc, pairs, table, flags=0, device_count=1):
    bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
    res = lib.align_batch(scoring_to_params(algo, sc, flags=flags, device_count=device_count), bases, off1, off2, len1, len2, sym_class=table)
    for p, (a, b) in enumerate(pairs):
        o = orc.oracle_align(algo, sc, translate_for_oracle(table, a, 0), translate_for_oracle(table, b, 1))
        got = (int(res.score[p]), int(res.start_i[p]), int(res.start_j[p]), int(res.end_i[p]), int(res.end_j[p]))
        assert got == (o["score"], o["start_i"], o["start_j"], o["end_i"], o["end_j"]), (algo, p, a, b, got, o)
        assert np.array_equal(res.pair_ops(p), o["ops"]), (algo, p, a, b)


@pytest.mark.parametrize("algo,sc", [("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("lgotoh", S.affine(-3, -1, 1, -1)),
                                     ("hirschberg", S.linear(-1, 2, -1)), ("myersmiller", S.affine(-3, -1, 1, -1))])
def test_table_driven_symbol_equality(emu_lib, algo, sc):
    """seqa_batch_in.sym_class: equivalent to aligning the class-id strings with ==."""
    rng = np.random.default_rng(41)
    for name, table in class_table_cases():
        alphabet = {"dna-caseless-N-never": "ACGTacgtNnR", "purine-pyrimidine": "ACGTacgtN", "protein-groups": "AVLFWSTKRDEGPX",
                    "near-identity": "ACGT\xfe\xff#"}[name]
        few = algo in ("hirschberg", "myersmiller")  # the emulated recursion is slow
        pairs = random_pairs(rng, 6 if few else 24, 1, 60, alphabet) + random_pairs(rng, 1 if few else 4, 70, 120, alphabet, related=0.3)
        check_class_table(emu_lib, algo, sc, pairs, table)
    check_class_table(emu_lib, algo, sc, pairs, table, flags=capi.FLAG_OPS_2BIT, device_count=2)
    bad = np.zeros(256, dtype=np.uint8)
    bad[65] = 254
    with pytest.raises(capi.SeqaError):
        emu_lib.align_batch(scoring_to_params(algo, sc), *orc.batch_arrays([("ACGT", "ACGA")]), sym_class=bad)


@pytest.mark.parametrize("algo,sc", [
    ("sw", S.linear(-1, 1, -1)), ("nw", S.linear(-1, 2, -1)), ("nw", S.linear(-14, 75, -1)), ("sw", S.linear(-29, 30, -30)),
    ("nw", S.linear(-2, 3, -1, False)), ("ggotoh", S.affine(-10, -40, 20, -30)), ("lgotoh", S.affine(-5, -5, 12, -9)),
])
def test_long_pairs_global_strip_boundaries(emu_lib, algo, sc):
    """Pairs above 320 columns: the packed fill keeps its strip-boundary rows in global memory as 8-bit differences along
    the row (pk_fill_kernel<GB>); scorings at both ends of the difference window, AllowMismatch off, related pairs."""
    rng = np.random.default_rng(17)
    pairs = random_pairs(rng, 5, 321, 420) + random_pairs(rng, 3, 330, 400, related=0.2) + random_pairs(rng, 3, 40, 350, "AC") + \
        [("ACGT" * 90, "ACGT" * 90), ("A" * 340, "C" * 330), ("A" * 25, "ACGT" * 85)]
    check_batch_against_oracle(emu_lib, algo, sc, pairs)


def test_speculative_batch_facts(emu_lib, monkeypatch):
    """Large one-device batches start their first wave before the pass over the index arrays has finished (planned as if
    every pair were like pair 0); a batch that stops being uniform behind the first wave keeps that wave's results and sends
    the rest through an ordinary second call.  Both outcomes against the oracle, then every wire format against that."""
    monkeypatch.setenv("SEQA_SPEC_MIN_PAIRS", "1000")
    rng = np.random.default_rng(23)
    uniform = random_pairs(rng, 4200, 12, 12)
    # same slots per pair (len1 + len2 = 24), other shapes: the caller's ops buffer still fits the assumed layout, so the first
    # wave really starts on the assumption and the helper thread is the one that refutes it
    odd = [(a[:10], b + a[10:]) for a, b in random_pairs(rng, 900, 12, 12)]
    ragged = uniform[:2500] + odd + uniform[2500:3000]  # fails behind the first wave
    early = uniform[:700] + random_pairs(rng, 10, 3, 20) + uniform[700:3500]      # fails inside the first wave: full pass
    for algo, sc, pairs in (("sw", S.linear(-1, 1, -1), uniform), ("sw", S.linear(-1, 1, -1), ragged), ("nw", S.linear(-1, 2, -1), ragged[1000:]),
                            ("sw", S.linear(-1, 1, -1), early)):
        check_batch_against_oracle(emu_lib, algo, sc, pairs)
        bases, off1, off2, len1, len2 = orc.batch_arrays(pairs)
        ref = emu_lib.align_batch(scoring_to_params(algo, sc), bases, off1, off2, len1, len2)
        pk, p1, p2 = capi.pack_bases_2bit(bases, off1, off2, len1, len2)
        for flags, ins in ((capi.FLAG_OPS_2BIT, (bases, off1, off2)), (capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT, (pk, p1, p2))):
            got = emu_lib.align_batch(scoring_to_params(algo, sc, flags=flags), ins[0], ins[1], ins[2], len1, len2)
            for f in ("score", "start_i", "start_j", "end_i", "end_j", "ops_len"):
                assert np.array_equal(getattr(got, f), getattr(ref, f)), (algo, flags, f)
            for q in range(0, len(pairs), 7):
                assert np.array_equal(got.pair_ops(q), ref.pair_ops(q)), (algo, flags, q)
