#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02k}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-api > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for c in d['configs']:
    print(c['config'], 'gcups %.0f ms %.1f frac %.2f fill %.0f checked %s' % (c['gcups'], c['ms'], c['roofline_frac'], c['fill_gcups'], c['oracle_checked_per_rank']))
PY
timeout 300 python tests/bench_configs.py 1 "config2 NW" 2>&1 | tail -1 | cut -c1-400
