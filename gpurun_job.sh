# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
timeout 300 python bench.py --no-cpu-baseline --tunable tc_fuse_cross=0 > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
timeout 300 python bench.py --no-cpu-baseline --tunable tc_tma_f32=0 > gpurun_out/bench_nof32tma.json 2> gpurun_out/bench_nof32tma.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 219 -c 73 --csv --log-file gpurun_out/launches_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_new.json | cut -c1-300; tail -3 gpurun_out/bench_new.err
