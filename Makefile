# Builds the product library (nvcc, sm_100a only) and the test-only checkers.
#   make            -> seqalib_b200/libseqa_cuda.so  (+ oracle checkers)
#   make emu        -> tests/emu/libseqa_emu.so      (same kernel sources on a CPU SIMT emulator; tests only)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
CSRC = seqalib_b200/csrc
HDRS = $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.inl) include/seqa_cuda.h
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-Wall -shared

all: seqalib_b200/libseqa_cuda.so oracle

seqalib_b200/libseqa_cuda.so: $(CSRC)/seqa_cuda.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v $< -o $@ 2> build_ptxas.log || (cat build_ptxas.log; false)

emu: tests/emu/libseqa_emu.so
tests/emu/libseqa_emu.so: $(CSRC)/seqa_cuda.cu $(HDRS) tests/emu/cuda_emu.h tests/emu/cuda_emu.cpp
	$(CXX) -std=c++17 -O1 -g -fPIC -shared -pthread -DSEQA_EMU -Itests/emu -I$(CSRC) -Wall -Wno-unused-function -Wno-unknown-pragmas \
	    -x c++ $(CSRC)/seqa_cuda.cu tests/emu/cuda_emu.cpp -o $@

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f seqalib_b200/libseqa_cuda.so tests/emu/libseqa_emu.so build_ptxas.log
	$(MAKE) -s -C oracle clean
.PHONY: all emu oracle clean
