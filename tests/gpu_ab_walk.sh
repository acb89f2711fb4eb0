#!/bin/bash
# A/B of the packed walk kernels on one B200: SEQA_WALK2=0 (pk_walk_kernel) vs 1 (pk_walk2_kernel)
mkdir -p gpurun_out
for w in ${AB_LIST:-0 1}; do
  echo "== SEQA_WALK2=$w"
  SEQA_WALK2=$w python bench.py --no-cpu --no-api --no-configs --steps 10 --warmup 3 2>gpurun_out/ab_err_$w.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f ms %.3f e2e %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
  SEQA_WALK2=$w python tests/bench_configs.py 1 "config2 NW" | grep -o '"gcups_step": [0-9.]*'
  SEQA_WALK2=$w python tests/bench_configs.py 1 "config5" | grep -o '"config": "[^"]*"\|"gcups_step": [0-9.]*' | paste - -
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or config2 or mixed or random" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:walk -c 2 --csv --log-file gpurun_out/ab_launches.csv python bench.py --no-cpu --no-api --no-configs --steps 1 --warmup 1 > /dev/null 2>&1
python tests/klist.py gpurun_out/ab_launches.csv
if [ -n "$AB_NCU" ]; then
ncu --set full --import-source on --clock-control none -k regex:walk2 -c 1 -o gpurun_out/walk2 -f python bench.py --no-cpu --no-api --no-configs --steps 1 --warmup 1 > /dev/null 2>&1
fi
