set -x
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -q > gpurun_out/r2_t32_full.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r2_t32_full.log
timeout 240 python bench.py > gpurun_out/r2_b32.json 2> gpurun_out/r2_b32.err; echo "bench rc=$?"
timeout 200 python bench.py --workload multihot --no-cpu-baseline > gpurun_out/r2_b32_multihot.json 2> gpurun_out/r2_b32_multihot.err; echo "bench rc=$?"
timeout 200 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b32_wide.json 2> gpurun_out/r2_b32_wide.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/r2_launches_v2.csv python bench.py --steps 20 --warmup 30 --no-cpu-baseline > gpurun_out/r2_ncu_b.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches_v2.csv | tail -50
python - <<'PY'
import json
for f in ("r2_b32","r2_b32_multihot","r2_b32_wide"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d.get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
