for lib in seqalib_b200/libseqa_cuda.so build_ab/libseqa_g2.so build_ab/libseqa_g3.so build_ab/libseqa_g5.so; do
echo "=== $lib"
SEQA_LIB=$PWD/$lib python tests/bench_configs.py 1 "config3 Global" | grep -o '"gcups_step": [0-9.]*\|"gcups_fill": [0-9.]*' | paste - -
SEQA_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:walk -c 1 --csv --log-file gpurun_out/gx.csv python tests/bench_configs.py 1 "config3 Global" > /dev/null 2>&1
grep -o '"gpu__time_duration.sum","ns","[0-9,]*"' gpurun_out/gx.csv
done
