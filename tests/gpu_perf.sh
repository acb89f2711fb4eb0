#!/bin/bash
# Perf iteration on one B200: quick parity, bench, launch list, one ncu --set full capture of the fill (+ walk) kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-config2_shape or matrix_algorithms_random}" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:---no-cpu} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/ncu1.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pk_fill|pk_walk" -s 2 -c 2 -o gpurun_out/prof_pk -f python bench.py --steps 2 --warmup 1 --no-cpu --pairs 200000 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | head -30
