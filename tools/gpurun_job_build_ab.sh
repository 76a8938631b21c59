# same-job A/B of two builds of the library (tools/_prev/libusot_b200_prev.so vs the current one), alternating
mkdir -p gpurun_out
for i in 1 2 3; do
USOT_B200_LIB=$PWD/tools/_prev/libusot_b200_prev.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/build_ab_prev_$i.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/build_ab_new_$i.json 2>/dev/null
done
for f in gpurun_out/build_ab_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1][-14:], round(d["value"]), round(d["e2e"]["value"]), round(d["kernel_ms_per_step"]["conv"],3), d["clocks"]["sm_mhz"])
PY
done
