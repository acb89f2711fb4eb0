for w in 0 1; do
SEQA_WALK2=$w ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:walk -c 1 --csv --log-file gpurun_out/lgw_$w.csv python tests/bench_configs.py 1 "config3 Local" > /dev/null 2>&1
grep -o 'pkg_walk[2]*_kernel<[^>]*>\|"gpu__time_duration.sum","ns","[0-9,]*"\|"dram__bytes_read.sum","[A-Za-z]*","[0-9.,]*"\|"smsp__inst_executed.sum","inst","[0-9,]*"\|"smsp__issue_active[^"]*","%","[0-9.]*"' gpurun_out/lgw_$w.csv | tr '\n' ' '; echo
done
