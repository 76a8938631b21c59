mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu.json')); print('8gpu value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tests/bench_cycle.py 2>&1 | tail -1 | tee gpurun_out/cycle_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_4gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu.json')); print('4gpu value', d['value'], 'e2e', d['e2e']['value'])"
