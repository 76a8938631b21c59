# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python tests/bench_sweep.py --batches 1,8 --nq 7 --ours-only > gpurun_out/sweep_lat1.jsonl 2> gpurun_out/sweep_lat1.err
timeout 300 python tests/bench_sweep.py --batches 1,8 --nq 7 --ours-only --tunable groupdw_row_split=0 > gpurun_out/sweep_lat0.jsonl 2> gpurun_out/sweep_lat0.err
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
timeout 300 python tools/tracker_fps.py fp16 300 >> gpurun_out/tracker_fps.json 2>> gpurun_out/tracker_fps.err
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep_lat1.jsonl gpurun_out/sweep_lat0.jsonl | cut -c1-250; cat gpurun_out/tracker_fps.json
