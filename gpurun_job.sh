# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
timeout 300 python bench.py --no-cpu-baseline --precision fp16 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 400 python bench.py > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 219 -c 73 --csv --log-file gpurun_out/launches_fp16.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision fp16 > gpurun_out/launches_fp16.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/tracker_fps.json; cut -c1-200 gpurun_out/bench_fp16.json; cut -c1-200 gpurun_out/bench_new.json; cut -c1-300 gpurun_out/bench_reference.json
