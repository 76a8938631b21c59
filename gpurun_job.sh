# GPU job of the current iteration (run as: gpurun --timeout 900 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_crop.py tests/test_tracker.py tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_crop.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:crop_resize -s 2 -c 1 -o gpurun_out/r01b_crop_resize -f python tools/crop_case.py > gpurun_out/ncu_crop.log 2>&1
tail -4 gpurun_out/pytest_gpu_crop.log; tail -2 gpurun_out/ncu_crop.log
