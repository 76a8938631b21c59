# N-GPU evidence job (N = 4 or 8):  gpurun --gpus N --timeout 420 -- 'bash tools/gpurun_job_ngpu.sh N'     (the whole job is bounded: N x 7 min)
# default line (carries BOTH collective records), BASELINE config 4 at the stated size as forward, config 5 sweep, config-4 training step.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
timeout 90 $TR --master-port 29523 bench.py --gpus $N --config 4 --steps 10 --no-collective > gpurun_out/bench_${N}gpu_c4.json 2> gpurun_out/bench_${N}gpu_c4.err
timeout 120 $TR --master-port 29522 bench.py --gpus $N --config 5 > gpurun_out/bench_${N}gpu_c5.json 2> gpurun_out/bench_${N}gpu_c5.err
for f in bench_${N}gpu bench_${N}gpu_c4 bench_${N}gpu_c5; do echo "== $f"; tail -1 gpurun_out/$f.json | cut -c1-260; tail -2 gpurun_out/$f.err | cut -c1-200; done
python - <<P
import json
for f in ("bench_${N}gpu","bench_${N}gpu_c4"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], json.dumps(d.get("collective"))[:1300], json.dumps(d.get("collective_train"))[:900])
    except Exception as e: print(f, "no line", e)
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_c5.json").read().strip().splitlines()[-1]); print(json.dumps(d["sweep"]))
except Exception as e: print("c5", e)
P
