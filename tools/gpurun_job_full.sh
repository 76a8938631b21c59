# full GPU validation of the current build: pytest -m gpu, smoke, default bench lines, launch list + one --set full capture of the pair kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -12 > gpurun_out/r02c_pytest_gpu.log
tail -4 gpurun_out/r02c_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02c_smoke.log
timeout 400 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/r02c_bench_fp16.json 2> gpurun_out/r02c_bench_fp16.err
timeout 300 python bench.py --no-cpu-baseline --config 3 > gpurun_out/r02c_bench_c3.json 2> gpurun_out/r02c_bench_c3.err
timeout 300 python bench.py --no-cpu-baseline --config 4 > gpurun_out/r02c_bench_c4.json 2> gpurun_out/r02c_bench_c4.err
timeout 300 python bench.py --no-cpu-baseline --config 4 --train-step > gpurun_out/r02c_bench_c4_train.json 2> gpurun_out/r02c_bench_c4_train.err
timeout 400 python bench.py --impl reference > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err
for f in r02c_bench r02c_bench_fp16 r02c_bench_c3 r02c_bench_c4 r02c_bench_c4_train r02c_bench_reference; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["e2e"]["value"],1) if "e2e" in d else None, round(d.get("ms_per_step",0),3), d.get("clocks",{}).get("sm_mhz"), {k: round(v,3) for k,v in d.get("kernel_ms_per_step",{}).items()}, round(d["roofline"]["frac"],4) if "roofline" in d else None, d.get("gpu_launches"))
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/r02c_launches_fp16x3_b256_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_launches_bench.log 2>&1
timeout 300 env USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02c_conv_tc_pair_l3_down -f python tools/conv_cases.py l3_down > gpurun_out/ncu_pair_a.log 2>&1
timeout 300 env USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02c_conv_tc_pair_l3_conv2 -f python tools/conv_cases.py l3_conv2 > gpurun_out/ncu_pair_b.log 2>&1
ls -la gpurun_out/r02c_*.ncu-rep; wc -l gpurun_out/r02c_launches_fp16x3_b256_final.csv
