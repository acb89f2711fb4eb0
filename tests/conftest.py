import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)["vectors"]


@pytest.fixture(scope="session")
def oracle_built():
    """The C restatement is test infrastructure; (re)build it when stale."""
    from oracle import pyoracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def emu_lib(oracle_built):
    """The product's kernel sources compiled for the CPU SIMT emulator (tests/emu) -- test-only."""
    subprocess.check_call(["make", "-s", "-C", ROOT, "emu"])
    from seqalib_b200 import capi
    return capi.Lib(os.path.join(ROOT, "tests", "emu", "libseqa_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib(oracle_built):
    """The real library on a real device; fails loudly (no fallback) when either is missing."""
    from seqalib_b200 import capi
    lib = capi.Lib()
    assert lib.device_count() >= 1, "no CUDA device visible"
    return lib
