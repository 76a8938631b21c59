# Full validation job on one B200 (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh'); tools/gpurun_job_2gpu.sh is the 2-GPU one.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
tail -5 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cut -c1-250 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_reference.json; cat gpurun_out/tracker_fps.json
