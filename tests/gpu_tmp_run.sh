python -m pytest tests/test_gpu_parity_edges.py -m gpu -x -q -k "speculative" 2>&1 | tail -8
