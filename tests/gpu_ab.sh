#!/bin/bash
# A/B of an experimental build of the library (SEQA_LIB) against the default one: parity subset, headline bench, launch list.
mkdir -p gpurun_out
for v in base ${VARIANTS:-pp}; do
  if [ $v = base ]; then unset SEQA_LIB; else export SEQA_LIB=$PWD/seqalib_b200/libseqa_cuda_$v.so; fi
  timeout 600 python -m pytest tests -m gpu -x -q -k "config2_shape or matrix_algorithms_random or mixed_length" > gpurun_out/ab_${v}_pytest.log 2>&1; echo "$v pytest rc=$?"; tail -2 gpurun_out/ab_${v}_pytest.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_${v}_bench.json 2> gpurun_out/ab_${v}_bench.err; echo "$v bench rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/ab_${v}_bench.json'));print('$v','value',round(d['value']),'ms',round(d['ms_per_step'],3),'fill ms',round(d['roofline']['kernel_ms_per_launch'],3),'e2e',round(d['e2e']['value']))"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/ab_${v}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ab_${v}_ncu.log 2>&1
  python tests/klist.py gpurun_out/ab_${v}_launches.csv | grep "pk_"
done
