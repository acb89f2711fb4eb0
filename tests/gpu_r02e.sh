#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02e}
echo "== two compute streams"; SWEEP_WAVES=2600,5100,7700 timeout 300 python tests/e2e_sweep.py 2>&1 | grep 2bit-in
echo "== one compute stream"; SEQA_ONE_COMPUTE_STREAM=1 SWEEP_WAVES=2600,5100,7700,10200 timeout 300 python tests/e2e_sweep.py 2>&1 | grep 2bit-in
echo "== one compute stream, ring 3/6"; SEQA_ONE_COMPUTE_STREAM=1 SWEEP_RINGS=3,6 SWEEP_WAVES=5100,7700 timeout 300 python tests/e2e_sweep.py 2>&1 | grep 2bit-in
SEQA_ONE_COMPUTE_STREAM=1 SEQA_WAVE_MCELLS=5100 timeout 300 python tests/e2e_probe.py > gpurun_out/${T}_e2e_probe_one_5100.txt 2>&1
grep -E "align_batch call|====|wave  " gpurun_out/${T}_e2e_probe_one_5100.txt | tail -12
