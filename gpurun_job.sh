# GPU job of the current iteration (run as: gpurun --timeout 1500 -- 'bash gpurun_job.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
timeout 300 python bench.py --no-cpu-baseline --tunable groupdw_tma=1 --tunable pred_tma_min_batch=0 > gpurun_out/bench_oldkernels.json 2> gpurun_out/bench_oldkernels.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:groupdw_ffma2 -c 2 -o gpurun_out/r01b_groupdw_ffma2 -f python tools/groupdw_case.py > gpurun_out/ncu_groupdw.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pred_gemm -c 2 -o gpurun_out/r01b_pred_gemm -f python tools/pred_case.py > gpurun_out/ncu_pred.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_new.json | cut -c1-1500; tail -3 gpurun_out/bench_new.err
