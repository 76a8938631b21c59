mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "memory or config1 or independence" 2>&1 | tail -3
run() {
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --csv --log-file gpurun_out/v_$1.csv env $2 python tools/conv_cases.py $4 256 $3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/v_$1.csv")) if len(r)>10 and r[0].isdigit()]
print("$1", "$2", "$3", "$4", [round(float(r[-1])/1e3,1) for r in rows])
PY
}
C=l3_conv3,l3_conv3_nores,l1_conv3,l1_conv3_nores
run splitout_x3 "USOT_DEBUG_SPLIT_OUT=2" fp16x3 $C
run splitout_16 "USOT_DEBUG_SPLIT_OUT=2" fp16 $C
for prec in fp16x3 fp16; do
python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$prec.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$prec.json"))
print("$prec value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d["clocks"])
PY
done
