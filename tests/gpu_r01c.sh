#!/bin/bash
# Round-1 re-entry measurements: e2e device timeline of the one-shot call, packed-affine CTAs/SM A/B (config 3).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c_gpu.txt
timeout 300 python tests/e2e_dbg.py 2 > gpurun_out/c_e2e_dbg.txt 2>&1; echo "e2e rc=$?"
for b in 2 3; do
  SEQA_PKG_BPS=$b timeout 600 python tests/bench_configs.py 1 config3 > gpurun_out/c_config3_bps$b.jsonl 2> gpurun_out/c_config3_bps$b.err; echo "cfg3 bps$b rc=$?"
done
cat gpurun_out/c_config3_bps*.jsonl
tail -40 gpurun_out/c_e2e_dbg.txt
