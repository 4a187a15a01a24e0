python -m pytest tests/test_gpu_parity_tile.py -m gpu -x -q 2>&1 | tail -6
