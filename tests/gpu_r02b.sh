#!/bin/bash
# Round 2, second GPU call: the parity tests that failed / were added since call a, and the bench with 2-bit inputs + API legs.
mkdir -p gpurun_out
T=${TAG:-r02b}
timeout 1500 python -m pytest tests/test_gpu_parity_edges.py tests/test_cpp_header.py -m gpu -x -q --durations=15 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-configs > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cat gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 300 python tests/e2e_probe.py > gpurun_out/${T}_e2e_probe.txt 2>&1; tail -40 gpurun_out/${T}_e2e_probe.txt
