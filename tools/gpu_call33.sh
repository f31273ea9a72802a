timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shards or pipelined or c1_single" 2>&1 | tail -3
