mkdir -p gpurun_out
PAIR_ALL=15 PAIR_SEL=11 timeout 300 python tools/pair_case.py ops fp16x3 > gpurun_out/pair_bn64_ops.log 2>&1; echo "rc=$?" >> gpurun_out/pair_bn64_ops.log; tail -15 gpurun_out/pair_bn64_ops.log
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/pair_launches_fp16x3_pairbn64.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tunable tc_cta_pair=15 > gpurun_out/pair_launches_bn64.log 2>&1
wc -l gpurun_out/pair_launches_fp16x3_pairbn64.csv
