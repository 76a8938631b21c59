# ncu evidence job (one B200): launch list of one timed step of the default bench + --set full captures of the top kernels.  Numbers printed under
# ncu are never bench values.   gpurun --timeout 1500 -- 'bash gpurun_job_ncu.sh'
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/r02_launches_fp16x3_b256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
timeout 300 env USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02_conv_tc_l3_down -f python tools/conv_cases.py l3_down > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:groupdw_ffma2 --launch-skip 2 -c 1 -o gpurun_out/r02_groupdw_ffma2_w4 -f python tools/groupdw_case.py > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02_wgrad_tc_l3_conv2 -f python tools/wgrad_case.py l3_conv2 > gpurun_out/ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02_wgrad_tc_l3_conv3 -f python tools/wgrad_case.py l3_conv3 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log; wc -l gpurun_out/r02_launches_fp16x3_b256.csv; tail -3 gpurun_out/r02_launches_bench.log | cut -c1-300
