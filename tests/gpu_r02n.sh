#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02n}
for L in "" _nocodes; do
  echo "=== lib$L"; SEQA_LIB=$PWD/seqalib_b200/libseqa_cuda$L.so timeout 600 python tests/bench_configs.py 1 "config" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        if 'config4' in d['config']: continue
        print('%-34s step %5.0f fill %5.0f GCUPS  ms %.1f  %s' % (d['config'], d['gcups_step'], d['gcups_fill'], d['ms_step'], d['kernel']))
    else: print(l.rstrip()[:200])"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_edges.py -m gpu -x -q -k "not config4 and not thousand and not linear_space" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
SEQA_API_TIMING=1 timeout 300 python bench.py --no-configs --no-cpu 2> gpurun_out/${T}_api.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.3f ms api_packed %s api_list %s' % (d['value'], d['e2e']['ms_per_step'], d['e2e']['api_packed']['gcups'], d['e2e']['api_list']['gcups']))"
g++ -std=c++14 -O2 -pthread -Iinclude tests/cpp/bench_header.cpp -o /tmp/bh -Lseqalib_b200 -lseqa_cuda -Wl,-rpath,$PWD/seqalib_b200 && SEQA_API_TIMING=1 /tmp/bh 1000000 3 200000 2>&1 | tail -8
