#!/bin/bash
mkdir -p gpurun_out
g++ -std=c++14 -O2 -pthread -Iinclude tests/cpp/bench_header.cpp -o /tmp/bh -Lseqalib_b200 -lseqa_cuda -Wl,-rpath,$PWD/seqalib_b200 && SEQA_API_TIMING=1 /tmp/bh 1000000 3 200000 2>&1 | tail -7 | cut -c1-420
timeout 300 python -m pytest tests/test_cpp_header.py -m gpu -x -q 2>&1 | tail -2
