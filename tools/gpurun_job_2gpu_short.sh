mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -4 gpurun_out/pytest_gpu_dist.log; cut -c1-400 gpurun_out/bench_2gpu.json; tail -2 gpurun_out/bench_2gpu.err
