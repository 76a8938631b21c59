mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -s -k "memory or config1 or independence or fresh" 2>&1 | grep -oE "(damp025|raw) [a-z0-9]+ fp16x3 \{[^}]*\}|backbone fresh [a-z0-9]+ [0-9.e-]+|[0-9]+ (passed|failed).*|Error.*|error.*" | tail -14
for prec in fp16x3 fp16; do
python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$prec.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$prec.json"))
print("$prec value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d["clocks"])
PY
done
