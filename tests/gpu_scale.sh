#!/bin/bash
# N-GPU run of bench.py exactly as the driver launches it (torchrun, one rank per GPU); N from $1
N=${1:-8}
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_${N}gpu.json'))
print('N=%d value %.0f ms %.3f e2e %.0f %.3f ms byte_bases %.3f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['byte_bases']['ms_per_step']))
for c in d['configs']:
    print(c['config'], 'gcups %.0f ms %.1f frac %.2f imbalance %.3f checked %s' % (c['gcups'], c['ms'], c['roofline_frac'], c['imbalance'], c['oracle_checked_per_rank']))
PY
