# A/B of tc_cta_pair at the other configs (same job, alternating)
mkdir -p gpurun_out
for i in 1 2; do
for pair in 0 3; do
timeout 300 python bench.py --no-cpu-baseline --config 4 --train-step --tunable tc_cta_pair=$pair > gpurun_out/ab2_c4train_p${pair}_$i.json 2> gpurun_out/ab2_c4train_p${pair}_$i.err
timeout 300 python bench.py --no-cpu-baseline --config 4 --tunable tc_cta_pair=$pair > gpurun_out/ab2_c4_p${pair}_$i.json 2> gpurun_out/ab2_c4_p${pair}_$i.err
done
done
for pair in 0 3; do
timeout 400 python bench.py --no-cpu-baseline --config 5 --tunable tc_cta_pair=$pair > gpurun_out/ab2_c5_p${pair}.json 2> gpurun_out/ab2_c5_p${pair}.err
done
for f in gpurun_out/ab2_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    for line in open(sys.argv[1]).read().strip().splitlines():
        d=json.loads(line)
        sw = d.get("sweep") or d.get("config",{}).get("sweep")
        print(round(d["value"],1), round(d.get("ms_per_step",0),3), d.get("clocks",{}).get("sm_mhz"), (json.dumps(sw)[:600] if sw else ""))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1][:-4]+"err").read()[-800:])
PY
done
