# A/B of the CTA-pair conv kernels (tc_cta_pair) inside the default bench; bit-identity first
mkdir -p gpurun_out
for prec in fp16x3 fp16; do
  timeout 400 python tools/pair_case.py model $prec > gpurun_out/pair_model_$prec.log 2>&1; echo "rc=$?" >> gpurun_out/pair_model_$prec.log
  echo "== model $prec"; tail -9 gpurun_out/pair_model_$prec.log
done
for i in 1 2; do
for pair in 0 3; do
timeout 300 python bench.py --no-cpu-baseline --tunable tc_cta_pair=$pair > gpurun_out/pair_bench_p${pair}_$i.json 2> gpurun_out/pair_bench_p${pair}_$i.err
done
done
for pair in 0 3; do
timeout 300 python bench.py --no-cpu-baseline --precision fp16 --tunable tc_cta_pair=$pair > gpurun_out/pair_bench_fp16_p${pair}.json 2> gpurun_out/pair_bench_fp16_p${pair}.err
timeout 300 python bench.py --no-cpu-baseline --config 3 --tunable tc_cta_pair=$pair > gpurun_out/pair_bench_c3_p${pair}.json 2> gpurun_out/pair_bench_c3_p${pair}.err
done
for f in pair_bench_p0_1 pair_bench_p3_1 pair_bench_p0_2 pair_bench_p3_2 pair_bench_fp16_p0 pair_bench_fp16_p3 pair_bench_c3_p0 pair_bench_c3_p3; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]) if "e2e" in d else None, round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d.get("kernel_ms_per_step",{}).items()}, round(d["roofline"]["frac"],4) if "roofline" in d else None)
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
