# Second-session job of round 2: new fused kernels (Conf_Fusion epilogue, stem+maxpool as a TMA-fed GEMM) -- parity tests first, then A/B bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_standalone_ops.py tests/test_gpu_tunables.py -q -m gpu -x --durations=5 -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2b_pytest.log
tail -45 gpurun_out/r2b_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python bench.py --no-cpu-baseline --tunable stem_pool_fused=0 > gpurun_out/r2b_bench_nopoolfuse.json 2> gpurun_out/r2b_bench_nopoolfuse.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/r2b_bench_fp16.json 2> gpurun_out/r2b_bench_fp16.err
for f in r2b_bench r2b_bench_nopoolfuse r2b_bench_fp16; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d["kernel_ms_per_step"].items()}, round(d["roofline"]["frac"],4), round(d["xcorr_roofline"]["frac"],3))
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
