#!/bin/bash
# One perf iteration on a B200: parity subset, headline bench (device + e2e), per-kernel launch list.
mkdir -p gpurun_out
T=${TAG:-iter}
timeout 900 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-golden or matrix_algorithms_random or config2_shape or mixed_length}" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python tests/klist.py gpurun_out/${T}_launches.csv 2>/dev/null | tail -20
