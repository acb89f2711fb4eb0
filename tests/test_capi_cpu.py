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
    assert lib.L.seqa_cuda_abi_version() == 1


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
    prm = capi.make_params("lgotoh", gap_open=-3, gap_extend=-1, match=1, mismatch=-1)
    b2 = orc.batch_arrays([("A" * 60, "C" * 57)])
    with pytest.raises(capi.SeqaError) as e:  # reference UB shape, include/SALocalGotoh.h:484-488
        emu_lib.align_batch(prm, *b2)
    assert e.value.code == -2
    small = capi.Results(1, 2)
    with pytest.raises(capi.SeqaError) as e:
        emu_lib.align_batch(scoring_to_params("nw", orc.Scoring.linear(-1, 2)), bases, off1, off2, len1, len2, small)
    assert e.value.code == -5
