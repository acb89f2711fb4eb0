#!/bin/bash
# Evidence run: launch list + ncu --set full of the headline kernels, config-3 fill, then plain bench lines.
mkdir -p gpurun_out
T=${TAG:-r02m}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pk_fill|pk_walk|pk_prep" -c 3 -o gpurun_out/${T}_prof_pk -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu pk rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pkg_fill|pkg_walk" -s 2 -c 2 -o gpurun_out/${T}_prof_pkg -f python tests/bench_configs.py 1 "config3 GlobalGotoh" > gpurun_out/${T}_ncu3.log 2>&1; echo "ncu pkg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ls_sweep2" -s 2 -c 1 -o gpurun_out/${T}_prof_ls_hb -f python tests/bench_configs.py 1 "config4 Hirschberg 100kbp x64" > gpurun_out/${T}_ncu4.log 2>&1; echo "ncu ls rc=$?"
timeout 300 python tests/int_peak.py > gpurun_out/${T}_int_peak.txt 2>&1
ls -la gpurun_out | grep ${T}_
