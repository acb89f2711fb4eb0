#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (old + new parity tests) and the bench line with the secondary configs.
mkdir -p gpurun_out
T=${TAG:-r02a}
nvidia-smi -L > gpurun_out/${T}_gpus.txt; nproc >> gpurun_out/${T}_gpus.txt; free -g >> gpurun_out/${T}_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
