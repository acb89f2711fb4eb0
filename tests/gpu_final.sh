#!/bin/bash
# End-of-round captures on one B200: smoke, headline bench (with cpu baseline), the reference arm, the launch list of the
# bench command, ncu --set full of its prep / fill / walk kernels.
mkdir -p gpurun_out
T=${TAG:-fin}
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cat gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/${T}_bench_reference.json
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pk_fill|pk_walk|pk_prep" -c 3 -o gpurun_out/${T}_prof_pk -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | grep ${T}_
