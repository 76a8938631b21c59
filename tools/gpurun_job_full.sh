# full GPU validation of the current build: pytest -m gpu, smoke, default bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -12 > gpurun_out/r02c_pytest_gpu.log
tail -12 gpurun_out/r02c_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02c_smoke.log
timeout 400 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/r02c_bench_fp16.json 2> gpurun_out/r02c_bench_fp16.err
timeout 300 python bench.py --no-cpu-baseline --config 3 > gpurun_out/r02c_bench_c3.json 2> gpurun_out/r02c_bench_c3.err
for f in r02c_bench r02c_bench_fp16 r02c_bench_c3; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]) if "e2e" in d else None, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d.get("kernel_ms_per_step",{}).items()}, round(d["roofline"]["frac"],4) if "roofline" in d else None, d.get("gpu_launches"))
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
