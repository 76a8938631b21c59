# N-GPU evidence job (N = 4 or 8):  gpurun --gpus N --timeout 1500 -- 'bash tools/gpurun_job_ngpu.sh N'
# BASELINE config 4 at the stated size (8 GPUs x 16 samples = global batch 128) as forward and as a training step, config 5 sweep, default line.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
timeout 500 $TR --master-port 29522 bench.py --gpus $N --config 5 > gpurun_out/bench_${N}gpu_c5.json 2> gpurun_out/bench_${N}gpu_c5.err
timeout 500 $TR --master-port 29523 bench.py --gpus $N --config 4 --steps 10 > gpurun_out/bench_${N}gpu_c4.json 2> gpurun_out/bench_${N}gpu_c4.err
timeout 500 $TR --master-port 29524 bench.py --gpus $N --config 4 --train-step --steps 5 > gpurun_out/bench_${N}gpu_train.json 2> gpurun_out/bench_${N}gpu_train.err
for f in bench_${N}gpu bench_${N}gpu_c5 bench_${N}gpu_c4 bench_${N}gpu_train; do echo "== $f"; tail -1 gpurun_out/$f.json | cut -c1-260; tail -2 gpurun_out/$f.err | cut -c1-200; done
python - <<P
import json
for f in ("bench_${N}gpu","bench_${N}gpu_c4","bench_${N}gpu_train"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], json.dumps(d.get("collective"))[:1200])
    except Exception as e: print(f, "no line", e)
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_c5.json").read().strip().splitlines()[-1]); print(json.dumps(d["sweep"]))
except Exception as e: print("c5", e)
P
