#!/bin/bash
# First GPU call of the round: parity tests, bench, launch list + one ncu capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tests/int_peak.py > gpurun_out/int_peak.txt 2>&1; cat gpurun_out/int_peak.txt
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/ncu1.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pk_fill -s 1 -c 1 -o gpurun_out/prof_pkfill -f python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
