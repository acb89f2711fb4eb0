#!/bin/bash
g++ -std=c++14 -O2 -pthread -Iinclude tests/cpp/bench_header.cpp -o /tmp/bh -Lseqalib_b200 -lseqa_cuda -Wl,-rpath,$PWD/seqalib_b200 && SEQA_DEBUG_TIMING=1 SEQA_API_TIMING=1 /tmp/bh 1000000 1 0 2>&1 | grep -v "upload 0\|^\[seqa\]   upload" | tail -34 | cut -c1-200
