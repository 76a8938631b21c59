# Full validation job on one B200 (run as: gpurun --timeout 2400 -- 'bash gpurun_job.sh'); tools/gpurun_job_2gpu.sh is the 2-GPU one,
# tools/gpurun_job_ngpu.sh the 4/8-GPU one, gpurun_job_ncu.sh the ncu evidence job.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 300 python bench.py --config 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 300 python bench.py --config 4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 400 python bench.py --config 4 --train-step --steps 5 > gpurun_out/bench_c4_train.json 2> gpurun_out/bench_c4_train.err
timeout 300 python tools/train_step_profile.py --graph > gpurun_out/train_profile_graph.json 2> gpurun_out/train_profile_graph.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/tracker_fps.py fp16x3 300 > gpurun_out/tracker_fps.json 2> gpurun_out/tracker_fps.err
tail -25 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; for f in bench bench_fp16 bench_c3 bench_c4 bench_c4_train bench_reference; do echo "== $f"; cut -c1-330 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done; cat gpurun_out/tracker_fps.json; head -5 gpurun_out/train_profile_graph.json
