python tests/bench_configs.py 1 "config5" | grep -o '"config": "[^"]*"\|"gcups_step": [0-9.]*\|"gcups_fill": [0-9.]*' | paste - - -
python -m pytest tests -m gpu -x -q -k "mixed or edge or thresh or 320 or long or config5" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fill -c 2 --csv --log-file gpurun_out/gb.csv python tests/bench_configs.py 1 "config5 mixed 50-1000bp SW" > /dev/null 2>&1
python tests/klist.py gpurun_out/gb.csv
