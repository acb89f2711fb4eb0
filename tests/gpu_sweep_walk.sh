#!/bin/bash
# sweep of the STEP-phase exit threshold of pk_walk2_kernel (SEQA_WALK2_T) on one B200
mkdir -p gpurun_out
for t in ${SWEEP:-0 2 4 6 8}; do
  echo "== SEQA_WALK2_T=$t"
  SEQA_WALK2_T=$t python bench.py --no-cpu --no-api --no-configs --steps 10 --warmup 3 2>gpurun_out/sw_err_$t.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f ms %.3f e2e %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
  SEQA_WALK2_T=$t python tests/bench_configs.py 1 "config2 NW" | grep -o '"gcups_step": [0-9.]*'
  SEQA_WALK2_T=$t python tests/bench_configs.py 1 "config5" | grep -o '"config": "[^"]*"\|"gcups_step": [0-9.]*' | paste - -
  SEQA_WALK2_T=$t ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum --clock-control none -k regex:walk -c 1 --csv --log-file gpurun_out/sw_launches_$t.csv python bench.py --no-cpu --no-api --no-configs --steps 1 --warmup 1 > /dev/null 2>&1
  grep -o '"gpu__time_duration.sum","ns","[0-9,]*"\|"dram__bytes_read.sum","[A-Za-z]*","[0-9.,]*"\|"smsp__inst_executed.sum","inst","[0-9,]*"' gpurun_out/sw_launches_$t.csv | tr '\n' ' '; echo
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or config2 or mixed or random" 2>&1 | tail -3
