# ncu launch list of ONE timed step of the default bench (final build of the second session of round 2)
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/r02c_launches_fp16x3_b256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_launches_bench.log 2>&1
wc -l gpurun_out/r02c_launches_fp16x3_b256.csv
