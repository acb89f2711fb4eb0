"""The 4-bit trace window of the packed affine kernels: the bounds the host relies on (packed_affine_trace_bits in
seqalib_b200/csrc/seqa_cuda.cu) hold on filled Gotoh matrices for every admitted scoring (numpy, no GPU)."""
import affine_window_check


def test_four_bit_window_bounds_hold():
    affine_window_check.main()
