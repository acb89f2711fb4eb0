#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02i}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pk_walk|pk_fill|pk_prep" -c 3 -o gpurun_out/${T}_prof_pk -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_ncu_pk.log 2>&1; echo "ncu pk rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ls_sweep2" -s 2 -c 1 -o gpurun_out/${T}_prof_ls -f python tests/bench_configs.py 1 "Hirschberg 100kbp x8" > gpurun_out/${T}_ncu_ls.log 2>&1; echo "ncu ls rc=$?"
ls -la gpurun_out/*.ncu-rep
