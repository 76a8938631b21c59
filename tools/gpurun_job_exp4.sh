mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/packed_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/packed_launches.log 2>&1
wc -l gpurun_out/packed_launches.csv
