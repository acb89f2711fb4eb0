mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9 --csv --log-file gpurun_out/l_sw.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python tests/klist.py gpurun_out/l_sw.csv | grep -E "walk|fill|prep|gather"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/l_nw.csv python tests/bench_configs.py 0.25 "config2 NW" > /dev/null 2>&1
python tests/klist.py gpurun_out/l_nw.csv | grep -E "walk|fill|prep|gather"
