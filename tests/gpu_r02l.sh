#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02l}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_edges.py -m gpu -x -q -k "not config4 and not thousand and not linear_space" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-configs --no-cpu --no-api > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench.json'))
r=d['roofline']
print('value %.0f ms %.3f | fill %.3f ms %.0f GCUPS alu %.2f | e2e %.3f ms' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['kernel_gcups'], r['alu_pipe']['frac'], d['e2e']['ms_per_step']), r['kernel'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/r02l_launches.csv') if not l.startswith('==')]
r=csv.reader(rows); h=next(r); ki,vi=h.index('Kernel Name'),h.index('Metric Value')
agg=collections.OrderedDict()
for row in r:
    if len(row)>vi: agg.setdefault(row[ki].split('(')[0],[]).append(float(row[vi].replace(',','')))
for k,v in agg.items(): print('%-40s n=%2d avg %.1f us' % (k,len(v),sum(v)/len(v)/1e3))
PY
timeout 300 python tests/bench_configs.py 1 "config2 NW" 2>&1 | tail -1 | cut -c1-330
echo "== TRACE4 for comparison"; 
