#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02j}
echo "== new walk"; timeout 300 python bench.py --no-configs --no-cpu --no-api | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f ms %.3f e2e %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
echo "== lazy walk"; SEQA_WALK_LAZY=1 timeout 300 python bench.py --no-configs --no-cpu --no-api | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f ms %.3f e2e %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/r02j_launches.csv') if not l.startswith('==')]
r=csv.reader(rows); h=next(r); ki,vi=h.index('Kernel Name'),h.index('Metric Value')
agg=collections.OrderedDict()
for row in r:
    if len(row)>vi: agg.setdefault(row[ki].split('(')[0],[]).append(float(row[vi].replace(',','')))
for k,v in agg.items(): print('%-40s n=%2d avg %.1f us' % (k,len(v),sum(v)/len(v)/1e3))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or matrix or config2 or mixed_length or two_bit" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
