mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tunables.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python tests/bench_sweep.py --batches 1,8 2>&1 | tail -2
timeout 600 python tests/bench_sweep.py --batches 1 --nq 7 2>&1 | tail -1
