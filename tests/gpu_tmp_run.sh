for lib in build_ab/libseqa_p3.so build_ab/libseqa_p5.so; do
echo "=== $lib"
SEQA_LIB=$PWD/$lib SWEEP="4" bash tests/gpu_sweep_walk.sh 2>&1 | grep -v passed
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
