# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
USOT_B200_TRACK_FRAME=0 timeout 300 python tools/tracker_fps.py fp16x3 300 >> gpurun_out/tracker_fps.json 2>> gpurun_out/tracker_fps.err
M=gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -s 219 -c 73 --csv --log-file gpurun_out/launches_pf1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tunable tc_l2_prefetch=1 > gpurun_out/launches_pf1.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 219 -c 73 --csv --log-file gpurun_out/launches_pf0.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_pf0.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 219 -c 73 --csv --log-file gpurun_out/launches_fp16_pf1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision fp16 --tunable tc_l2_prefetch=1 > gpurun_out/launches_fp16_pf1.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/tracker_fps.json
