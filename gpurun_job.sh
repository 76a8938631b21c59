timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "prroi" 2>&1 | tail -6
