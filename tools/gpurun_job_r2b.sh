# Second-session job of round 2: parity tests of the conv kernels, then A/B bench lines.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tunables.py tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_configs.py -q -m gpu -x --durations=3 2>&1 | tail -6 > gpurun_out/r2b_pytest.log
tail -6 gpurun_out/r2b_pytest.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_$i.json 2> gpurun_out/r2b_bench_$i.err
timeout 300 python bench.py --no-cpu-baseline --tunable tc_res_ahead=1 > gpurun_out/r2b_bench_ra1_$i.json 2> gpurun_out/r2b_bench_ra1_$i.err
done
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/r2b_bench_fp16.json 2> gpurun_out/r2b_bench_fp16.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 --tunable tc_res_ahead=1 > gpurun_out/r2b_bench_fp16_ra1.json 2> gpurun_out/r2b_bench_fp16_ra1.err
for f in r2b_bench_1 r2b_bench_ra1_1 r2b_bench_2 r2b_bench_ra1_2 r2b_bench_fp16 r2b_bench_fp16_ra1; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d["kernel_ms_per_step"].items()}, round(d["roofline"]["frac"],4), round(d["xcorr_roofline"]["frac"],3))
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
