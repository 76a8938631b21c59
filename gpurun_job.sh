timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -s -x 2>&1 | tail -25
