#!/bin/bash
# End-of-round evidence on ONE B200 (scratch -> gpurun_out/<tag>_*; tests/collect_profiles.py turns it into profiles/).
# Plain runs first (pytest, smoke, bench), then the profiler passes of the same commands.
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=4 > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${tag}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; tail -c 400 gpurun_out/${tag}_bench_1gpu.json; echo
if [ -n "$FINAL_REF" ]; then python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; fi
SEQA_API_TIMING=1 python bench.py --no-cpu --no-configs --steps 2 --warmup 1 2> gpurun_out/${tag}_api_timing.txt > /dev/null
# launch list of the headline command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_sw150_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${tag}_ncu1.log 2>&1
# full captures: headline prep / fill / walk; GlobalGotoh fill / walk
ncu --set full --import-source on --clock-control none -k regex:'pk_prep|pk_fill|pk_walk' -c 3 -f -o gpurun_out/${tag}_prof_pk \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${tag}_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'pkg_fill|pkg_walk' -c 2 -f -o gpurun_out/${tag}_prof_pkg \
    python tests/bench_configs.py 1 "config3 GlobalGotoh" > gpurun_out/${tag}_ncu3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'pk_fill|pk_walk' -c 2 -f -o gpurun_out/${tag}_prof_gb \
    python tests/bench_configs.py 1 "config5 mixed 50-1000bp SW" > gpurun_out/${tag}_ncu4.log 2>&1
ls -la gpurun_out/${tag}_*
