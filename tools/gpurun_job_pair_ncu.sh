# per-launch comparison of the CTA-pair conv kernels (tc_cta_pair=3) with the one-CTA kernels: ncu launch lists of one step each
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
for prec in fp16x3 fp16; do
for pair in 0 3; do
timeout 400 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/pair_launches_${prec}_pair${pair}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision $prec --tunable tc_cta_pair=$pair > gpurun_out/pair_launches_${prec}_pair${pair}.log 2>&1
wc -l gpurun_out/pair_launches_${prec}_pair${pair}.csv
done
done
