# ncu evidence job (one B200): launch list of one timed step of the default bench + --set full captures of the kernels that changed in the
# second session of round 2.  Numbers printed under ncu are never bench values.   gpurun --timeout 1500 -- 'bash gpurun_job_ncu.sh'
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none --launch-skip 210 --launch-count 70 --csv --log-file gpurun_out/r02b_launches_fp16x3_b256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_launches_bench.log 2>&1
timeout 300 env USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02b_conv_tc_l3_down -f python tools/conv_cases.py l3_down > gpurun_out/ncu_a.log 2>&1
timeout 300 env USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02b_conv_tc_l3_conv3 -f python tools/conv_cases.py l3_conv3 > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02b_stem_gemm_pool -f python tools/stem_case.py > gpurun_out/ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_s2d_kernel --launch-skip 2 -c 1 -o gpurun_out/r02b_stem_s2d -f python tools/stem_case.py > gpurun_out/ncu_d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 1 -c 1 -o gpurun_out/r02b_wgrad_tc_l3_conv2 -f python tools/wgrad_case.py l3_conv2 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/r02b_*.ncu-rep; for f in a b c d e; do tail -n 1 gpurun_out/ncu_$f.log; done; wc -l gpurun_out/r02b_launches_fp16x3_b256.csv; tail -n 2 gpurun_out/r02b_launches_bench.log | cut -c1-300
