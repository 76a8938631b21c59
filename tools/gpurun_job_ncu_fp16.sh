mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,sm__cycles_elapsed.max.per_second
timeout 600 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/r02b_launches_fp16_b256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision fp16 > gpurun_out/r02b_launches_fp16_bench.log 2>&1
wc -l gpurun_out/r02b_launches_fp16_b256.csv
