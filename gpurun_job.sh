# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
timeout 300 python tools/tracker_fps.py fp16 300 >> gpurun_out/tracker_fps.json 2>> gpurun_out/tracker_fps.err
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/tracker_fps.json; tail -3 gpurun_out/tracker_fps.err
