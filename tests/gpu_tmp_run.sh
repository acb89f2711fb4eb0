for lib in seqalib_b200/libseqa_cuda.so build_ab/libseqa_minb5.so; do
echo "=== $lib"
export SEQA_LIB=$PWD/$lib
SWEEP="4" bash tests/gpu_sweep_walk.sh
done
