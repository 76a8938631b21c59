mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_tunables.py -q -m gpu -x -k "groupdw or memory or config1 or knob or fresh" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_fp16x3.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_fp16x3.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "xcorr", round(d["xcorr_roofline"]["achieved"]), d["kernel_ms_per_step"], d["clocks"]["sm_mhz"])
PY
