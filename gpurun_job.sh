# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tunables.py tests/test_gpu_model.py tests/test_tracker.py -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
USOT_B200_TUNABLES=tc_pdl=1 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu_pdl.log
timeout 300 python tests/bench_sweep.py --batches 1,8,256 --nq 7 --ours-only --tunable tc_pdl=1 > gpurun_out/sweep_pdl1.jsonl 2> gpurun_out/sweep_pdl1.err
timeout 300 python tests/bench_sweep.py --batches 1,8,256 --nq 7 --ours-only > gpurun_out/sweep_pdl0.jsonl 2> gpurun_out/sweep_pdl0.err
USOT_B200_TUNABLES=tc_pdl=1 timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps_pdl.json 2> gpurun_out/tracker_fps_pdl.err
timeout 300 python bench.py --no-cpu-baseline --tunable tc_pdl=1 > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err
tail -4 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu_pdl.log; cat gpurun_out/sweep_pdl1.jsonl gpurun_out/sweep_pdl0.jsonl | cut -c1-250; cat gpurun_out/tracker_fps_pdl.json; cut -c1-150 gpurun_out/bench_pdl1.json; cut -c1-150 gpurun_out/bench_pdl0.json; tail -2 gpurun_out/bench_pdl1.err
