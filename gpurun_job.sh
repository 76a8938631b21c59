mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -10
python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('e2e',d['e2e']); print('roofline',d['roofline']); print('xcorr',d['xcorr_roofline']); print('cpu',d['cpu_baseline']); print(d['kernel_ms_per_step'])"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; python -c "
import json; d=json.load(open('gpurun_out/bench_reference.json')); print('ref', d['value'], d['cpu_baseline'])"
python bench.py --precision fp16 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_fp16_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16_final.json')); print('fp16', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])"
