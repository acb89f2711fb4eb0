#!/bin/bash
# End-of-round evidence on one B200: smoke, the whole -m gpu suite, the bench line (with configs[], cpu baseline, API legs), the
# reference arm, the launch list of the bench command, ncu --set full of the headline / affine / linear-space kernels, the
# API phase timing.  Outputs under gpurun_out/${TAG}_*; summarised into profiles/ by tests/collect_profiles.py.
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_sw150_1M.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pk_fill|pk_walk|pk_prep" -c 3 -o gpurun_out/${T}_prof_pk -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --no-api > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu pk rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pkg_fill|pkg_walk" -s 2 -c 2 -o gpurun_out/${T}_prof_pkg -f python tests/bench_configs.py 1 "config3 GlobalGotoh" > gpurun_out/${T}_ncu3.log 2>&1; echo "ncu pkg rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ls_sweep2" -c 1 -o gpurun_out/${T}_prof_ls_hb -f python tests/bench_configs.py 1 "config4 Hirschberg 100kbp x64" > gpurun_out/${T}_ncu4.log 2>&1; echo "ncu ls hb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ls_sweep2" -c 1 -o gpurun_out/${T}_prof_ls_mm -f python tests/bench_configs.py 1 "config4 MyersMiller 100kbp x64" > gpurun_out/${T}_ncu5.log 2>&1; echo "ncu ls mm rc=$?"
g++ -std=c++14 -O2 -pthread -Iinclude tests/cpp/bench_header.cpp -o /tmp/bh -Lseqalib_b200 -lseqa_cuda -Wl,-rpath,$PWD/seqalib_b200 && SEQA_API_TIMING=1 /tmp/bh 1000000 3 200000 > gpurun_out/${T}_api_timing.txt 2>&1; tail -3 gpurun_out/${T}_api_timing.txt | cut -c1-300
timeout 300 python tests/e2e_probe.py > gpurun_out/${T}_e2e_probe_final.txt 2>&1
ls gpurun_out | grep ${T}_ | tr '\n' ' '
