"""seqalib_b200 -- B200-native (sm_100a) DP fill + traceback behind SeqALib's aligner API.

The product is the shared library `libseqa_cuda.so` (C ABI: include/seqa_cuda.h) and the C++ header-only host
mirror of the reference API (include/SequenceAlignment.h).  This Python package only binds the C ABI for tests and
benchmarks.  Importing `seqalib_b200.capi.Lib()` raises ImportError when the CUDA library has not been built.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
