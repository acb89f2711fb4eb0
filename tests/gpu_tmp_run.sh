python tests/bench_configs.py 1 "config3" | grep -o '"config": "[^"]*"\|"gcups_step": [0-9.]*\|"gcups_fill": [0-9.]*' | paste - - -
python -m pytest tests -m gpu -x -q -k "gotoh or golden or config3 or thresh or edge" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fill -c 2 --csv --log-file gpurun_out/gg.csv python tests/bench_configs.py 1 "config3 Global" > /dev/null 2>&1
python tests/klist.py gpurun_out/gg.csv
