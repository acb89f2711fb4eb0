#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02d}
for W in 5100 7700 12800; do
  SEQA_WAVE_MCELLS=$W timeout 300 python tests/e2e_probe.py > gpurun_out/${T}_e2e_probe_$W.txt 2>&1
  echo "=== wave $W"; grep -E "align_batch call|====|wave  |upload\+plan|wait" gpurun_out/${T}_e2e_probe_$W.txt | tail -32
done
