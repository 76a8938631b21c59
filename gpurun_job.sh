# Validation job on one B200 (run as: gpurun --timeout 2400 -- 'bash gpurun_job.sh'); tools/gpurun_job_2gpu.sh is the 2-GPU one.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_backward.py -q -m gpu -x 2>&1 | grep -v Warning | tail -40 > gpurun_out/pytest_train.log
timeout 300 python tools/train_step_profile.py > gpurun_out/train_profile_eager.json 2> gpurun_out/train_profile_eager.err
timeout 300 python tools/train_step_profile.py --graph > gpurun_out/train_profile_graph.json 2> gpurun_out/train_profile_graph.err
timeout 400 python bench.py --config 4 --train-step --steps 5 > gpurun_out/bench_c4_train.json 2> gpurun_out/bench_c4_train.err
timeout 400 python bench.py --config 4 --train-step --no-graph --steps 5 > gpurun_out/bench_c4_train_eager.json 2> gpurun_out/bench_c4_train_eager.err
tail -30 gpurun_out/pytest_train.log; for f in train_profile_eager train_profile_graph; do echo "== $f"; head -4 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done; for f in bench_c4_train bench_c4_train_eager; do echo "== $f"; cut -c1-300 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
