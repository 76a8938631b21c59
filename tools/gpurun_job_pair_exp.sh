# experiments on the CTA-pair kernels: fused cross-term pair layout (bit 4), 3-stage residual variant (bit 3)
mkdir -p gpurun_out
PAIR_ALL=23 PAIR_SEL=19 timeout 300 python tools/pair_case.py ops fp16x3 > gpurun_out/pair_fused_ops.log 2>&1; echo "rc=$?" >> gpurun_out/pair_fused_ops.log; tail -13 gpurun_out/pair_fused_ops.log
PAIR_ALL=23 PAIR_SEL=19 timeout 300 python tools/pair_case.py model fp16x3 > gpurun_out/pair_fused_model.log 2>&1; echo "rc=$?" >> gpurun_out/pair_fused_model.log; tail -9 gpurun_out/pair_fused_model.log
PAIR_ALL=15 PAIR_SEL=15 timeout 300 python tools/pair_case.py model fp16x3 > gpurun_out/pair_res3_model.log 2>&1; echo "rc=$?" >> gpurun_out/pair_res3_model.log; tail -9 gpurun_out/pair_res3_model.log
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
for pair in 19 15; do
timeout 400 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/pair_launches_fp16x3_pair${pair}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tunable tc_cta_pair=$pair > gpurun_out/pair_launches_fp16x3_pair${pair}.log 2>&1
wc -l gpurun_out/pair_launches_fp16x3_pair${pair}.csv
done
