# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_tunables.py tests/test_gpu_train.py -q -m gpu 2>&1 | tail -20 > gpurun_out/pytest_gpu_stem.log
timeout 400 python bench.py > gpurun_out/bench_stem.json 2> gpurun_out/bench_stem.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/bench_stem_fp16.json 2> gpurun_out/bench_stem_fp16.err
tail -4 gpurun_out/pytest_gpu_stem.log; python -c "
import json
for f in ('bench_stem','bench_stem_fp16'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['clocks'])"
