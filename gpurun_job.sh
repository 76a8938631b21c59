# Validation job on one B200 (run as: gpurun --timeout 2400 -- 'bash gpurun_job.sh'); tools/gpurun_job_2gpu.sh is the 2-GPU one.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_backward.py -q -m gpu -s 2>&1 | grep -v Warning | tail -150 > gpurun_out/pytest_train.log
timeout 600 python -m pytest tests/test_gpu_tunables.py tests/test_gpu_ops.py tests/test_gpu_packed.py tests/test_gpu_standalone_ops.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 400 python bench.py --config 4 --train-step --steps 5 > gpurun_out/bench_c4_train.json 2> gpurun_out/bench_c4_train.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --tunable groupdw_warps4=0 --no-cpu-baseline > gpurun_out/bench_gdw3.json 2> gpurun_out/bench_gdw3.err
grep -v "^wgrad\|^$" gpurun_out/pytest_train.log | tail -40; grep "^wgrad" gpurun_out/pytest_train.log | head -40; tail -8 gpurun_out/pytest_gpu.log; for f in bench_c4_train bench bench_gdw3; do echo "== $f"; cut -c1-300 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
