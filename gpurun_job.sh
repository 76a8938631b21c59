mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_2gpu.json | cut -c1-600
