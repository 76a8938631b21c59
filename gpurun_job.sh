# Final GPU job of round 1, second session (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/bench_final_fp16.json 2> gpurun_out/bench_final_fp16.err
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
timeout 300 python tools/tracker_fps.py fp16 300 >> gpurun_out/tracker_fps.json 2>> gpurun_out/tracker_fps.err
timeout 300 python tests/bench_sweep.py --batches 1,8 --nq 7 --ours-only > gpurun_out/sweep_small.jsonl 2> gpurun_out/sweep_small.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -s 210 -c 70 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1
USOT_DEBUG_SPLIT_OUT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/r01b_conv_tc_l3_down -f python tools/conv_cases.py l3_down > gpurun_out/ncu_conv1.log 2>&1
USOT_DEBUG_SPLIT_OUT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/r01b_conv_tc_l3_conv3 -f python tools/conv_cases.py l3_conv3 > gpurun_out/ncu_conv2.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cut -c1-250 gpurun_out/bench_final.json; cut -c1-150 gpurun_out/bench_final_fp16.json; cat gpurun_out/tracker_fps.json; cut -c1-200 gpurun_out/sweep_small.jsonl; tail -2 gpurun_out/ncu_conv1.log
