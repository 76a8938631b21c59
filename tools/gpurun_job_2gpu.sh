# 2-GPU job: NCCL tests + torchrun bench lines at N=2 (run as: gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpurun_job_2gpu.sh')
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --config 4 --train-step --no-graph --steps 5 > gpurun_out/bench_2gpu_train_eager.json 2> gpurun_out/bench_2gpu_train_eager.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --config 4 --train-step --steps 5 > gpurun_out/bench_2gpu_train_graph.json 2> gpurun_out/bench_2gpu_train_graph.err
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --config 5 > gpurun_out/bench_2gpu_c5.json 2> gpurun_out/bench_2gpu_c5.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
tail -6 gpurun_out/pytest_gpu_dist.log; for f in bench_2gpu bench_2gpu_train_eager bench_2gpu_train_graph bench_2gpu_c5 bench_2gpu_ref; do echo "== $f"; cut -c1-250 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
python - <<'P'
import json
for f in ("bench_2gpu","bench_2gpu_train_eager","bench_2gpu_train_graph"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, json.dumps(d.get("collective"))[:1500])
    except Exception as e: print(f, "no line", e)
P
