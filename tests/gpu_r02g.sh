#!/bin/bash
# 2-GPU call: the multi-device tests of the library's own split + bench.py under torchrun with the secondary configs
mkdir -p gpurun_out
T=${TAG:-r02g}
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_device" > gpurun_out/${T}_pytest_2gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_2gpu.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for c in d['configs']:
    print(c['config'], 'gcups %.0f ms %.1f frac %.2f imbalance %.3f checked %s per_rank_ms %s' % (c['gcups'], c['ms'], c['roofline_frac'], c['imbalance'], c['oracle_checked_per_rank'], [round(x,1) for x in c['per_rank_ms']]))
PY
