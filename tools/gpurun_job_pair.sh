# CTA-pair (cta_group::2) conv kernels: bit-identity vs the one-CTA kernels, then informal step timing (tools/pair_case.py).
mkdir -p gpurun_out
for prec in fp16x3 fp16; do
  timeout 300 python tools/pair_case.py ops $prec > gpurun_out/pair_ops_$prec.log 2>&1
  rc=$?; echo "rc=$rc" >> gpurun_out/pair_ops_$prec.log
  echo "== ops $prec"; tail -14 gpurun_out/pair_ops_$prec.log
  if [ $rc -eq 0 ]; then
    timeout 400 python tools/pair_case.py model $prec > gpurun_out/pair_model_$prec.log 2>&1
    echo "rc=$?" >> gpurun_out/pair_model_$prec.log
    echo "== model $prec"; tail -10 gpurun_out/pair_model_$prec.log
  fi
done
