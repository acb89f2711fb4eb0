"""include/SequenceAlignment.h (the C++ host mirror of the reference API) compiled and run: against the emulator
build on CPU boxes, against libseqa_cuda.so on the GPU box."""
import os
import subprocess

import pytest

from common import ROOT


def _build_and_run(libdir, libname, tmp_path):
    exe = str(tmp_path / "test_header")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_header.cpp"), "-o", exe,
                           "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout
    assert "AAA-GAATGCAT\n|||    | |||\nAAAC---T-CAT" in out.stdout  # reference README.md:34-37


def _bridge_vs_oracle(libdir, libname, tmp_path):
    """StaticFuncs::bridgeNWBatch on anchor windows (as MUMmer / BLAT cut them) against oracle_align("nw") on the
    substrings -- the C++ test links the C oracle (test infrastructure) beside the library under test."""
    exe = str(tmp_path / "test_bridge_oracle")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_bridge_oracle.cpp"), "-o", exe,
                           "-L", libdir, "-l" + libname, "-L", odir, "-lseqa_oracle", "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().startswith("OK"), out.stdout
    assert int(out.stdout.split()[1]) > 100


def test_bridge_windows_against_oracle_emulated(emu_lib, tmp_path):
    _bridge_vs_oracle(os.path.join(ROOT, "tests", "emu"), "seqa_emu", tmp_path)


@pytest.mark.gpu
def test_bridge_windows_against_oracle_on_gpu(gpu_lib, tmp_path):
    _bridge_vs_oracle(os.path.join(ROOT, "seqalib_b200"), "seqa_cuda", tmp_path)


def test_header_against_emulated_kernels(emu_lib, tmp_path):
    _build_and_run(os.path.join(ROOT, "tests", "emu"), "seqa_emu", tmp_path)


@pytest.mark.gpu
def test_header_on_gpu(gpu_lib, tmp_path):
    _build_and_run(os.path.join(ROOT, "seqalib_b200"), "seqa_cuda", tmp_path)
