# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
timeout 400 python tests/bench_sweep.py --batches 1,8,64,256 > gpurun_out/sweep_offline.jsonl 2> gpurun_out/sweep_offline.err
timeout 400 python tests/bench_sweep.py --batches 1,8,64 --nq 7 > gpurun_out/sweep_nq7.jsonl 2> gpurun_out/sweep_nq7.err
tail -6 gpurun_out/pytest_gpu.log; cut -c1-200 gpurun_out/bench_new.json; cat gpurun_out/tracker_fps.json; cat gpurun_out/sweep_offline.jsonl gpurun_out/sweep_nq7.jsonl | cut -c1-420
