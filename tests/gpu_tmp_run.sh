nproc; free -g | head -2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 2 --no-configs --no-api --no-cpu 2>gpurun_out/e8_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$1 value %.0f ms %.3f e2e %.0f ms %.3f bytebases %.3f' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['byte_bases']['ms_per_step']))"; }
run default
SEQA_NO_SPECULATIVE_FACTS=1 run nospec
SEQA_DEBUG_TIMING=1 run debug
grep "dev 0 wave\|host before\|call returns\|threads joined" gpurun_out/e8_debug.err | tail -40
