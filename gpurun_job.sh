mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s > gpurun_out/model_tests.log 2>&1
grep -oE "(damp025|raw) [a-z0-9]+ fp[0-9x]+ \{[^}]*\}|backbone fresh [a-z0-9]+ [0-9.e-]+|[0-9]+ (passed|failed).*" gpurun_out/model_tests.log
