# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_pf1.json 2> gpurun_out/bench_pf1.err
timeout 300 python bench.py --no-cpu-baseline --tunable tc_l2_prefetch=0 > gpurun_out/bench_pf0.json 2> gpurun_out/bench_pf0.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_pf1b.json 2> gpurun_out/bench_pf1b.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/bench_fp16_pf1.json 2> gpurun_out/bench_fp16_pf1.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 --tunable tc_l2_prefetch=0 > gpurun_out/bench_fp16_pf0.json 2> gpurun_out/bench_fp16_pf0.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 219 -c 73 --csv --log-file gpurun_out/launches_pf1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_pf1.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; for f in pf1 pf0 pf1b fp16_pf1 fp16_pf0; do cut -c1-130 gpurun_out/bench_$f.json; done
