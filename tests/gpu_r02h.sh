#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02h}
for L in "" _ls75 _ls85; do
  echo "=== lib$L"; SEQA_LIB=$PWD/seqalib_b200/libseqa_cuda$L.so timeout 600 python tests/bench_configs.py 1 "100kbp x64" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'step %.0f fill %.0f GCUPS  ms %.1f' % (d['gcups_step'], d['gcups_fill'], d['ms_step']))
    else: print(l.rstrip()[:200])"
done
echo "=== default lib, SEQA_LS_BPS=4 (round-1 grid)"; SEQA_LS_BPS=4 timeout 600 python tests/bench_configs.py 1 "100kbp x64" 2>&1 | grep -o '"config": "[^"]*"\|"gcups_step": [0-9.]*' | paste - - 
