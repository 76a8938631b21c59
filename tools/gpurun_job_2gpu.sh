# 2-GPU job: NCCL sharding test + torchrun bench line at N=2 (run as: gpurun --gpus 2 --timeout 900 -- 'bash tools/gpurun_job_2gpu.sh')
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
tail -4 gpurun_out/pytest_gpu_dist.log; cut -c1-300 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err; cut -c1-200 gpurun_out/bench_2gpu_ref.json
