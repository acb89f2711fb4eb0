#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02c}
timeout 600 python tests/e2e_sweep.py > gpurun_out/${T}_e2e_sweep.txt 2>&1; cat gpurun_out/${T}_e2e_sweep.txt | tail -20
timeout 300 python tests/e2e_probe.py > gpurun_out/${T}_e2e_probe.txt 2>&1; grep -E "align_batch call|====|wave  " gpurun_out/${T}_e2e_probe.txt | tail -40
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_bit or config2_shape or golden or multi_device or device_resident" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
