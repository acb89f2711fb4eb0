#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02f}
timeout 300 python tests/e2e_probe.py > gpurun_out/${T}_e2e_probe_default.txt 2>&1
grep -E "align_batch call|====|wave  " gpurun_out/${T}_e2e_probe_default.txt | tail -24
for R in 2 4; do echo "== SEQA_WAVE_ROUNDS=$R"; SEQA_WAVE_ROUNDS=$R timeout 300 python tests/e2e_probe.py 2>&1 | grep -E "align_batch call [123]|====" ; done
echo "== two compute streams"; SEQA_TWO_COMPUTE_STREAMS=1 timeout 300 python tests/e2e_probe.py 2>&1 | grep -E "align_batch call [123]|===="
timeout 600 python bench.py --no-configs --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print('value',d['value'],'ms',d['ms_per_step']); e=d['e2e']
print('e2e',e['value'],e['ms_per_step'],'byte_bases',e['byte_bases']['ms_per_step'],'byte_ops',e['byte_ops']['ms_per_step'])
print(e['api_packed']); print(e['api_list'])
PY
