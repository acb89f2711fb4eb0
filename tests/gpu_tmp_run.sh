python tests/e2e_probe.py 2>&1 | sed -n '/2-bit symbols in, 2-bit ops out/,$p' | grep -v "upload+plan\|wait "
python bench.py --no-cpu --no-configs --no-api --steps 10 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value %.0f ms %.3f e2e %.3f bytebases %.3f byteops %.3f' % (d['value'], d['ms_per_step'], e['ms_per_step'], e['byte_bases']['ms_per_step'], e['byte_ops']['ms_per_step']))"
python -m pytest tests -m gpu -x -q -k "wire or two_bit or 2bit or full_size or layout or wave" 2>&1 | tail -3
