# Full validation job on one B200 (run as: gpurun --timeout 2400 -- 'bash gpurun_job.sh'); tools/gpurun_job_2gpu.sh is the 2-GPU one.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_backward.py -q -m gpu -s 2>&1 | grep -v Warning | tail -60 > gpurun_out/pytest_train.log
timeout 1500 python -m pytest tests -q -m gpu --durations=10 --deselect tests/test_gpu_train_ops.py --deselect tests/test_gpu_train_backward.py 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --config 4 --train-step --steps 5 > gpurun_out/bench_c4_train.json 2> gpurun_out/bench_c4_train.err
timeout 400 python bench.py --config 4 --steps 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -45 gpurun_out/pytest_train.log; tail -15 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; for f in bench_c4_train bench_c4; do echo "== $f"; cut -c1-600 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
