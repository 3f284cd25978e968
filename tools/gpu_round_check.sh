# One GPU call that re-validates a round: full GPU suite, the three bench workloads, a step timeline and the ncu launch list.
#   tools/gpurun_retry.sh gpurun_out/check.log --timeout 1000 -- "bash tools/gpu_round_check.sh"
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2_t38_full.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r2_t38_full.log
timeout 200 python bench.py > gpurun_out/r2_b38.json 2> gpurun_out/r2_b38.err; echo "bench rc=$?"
timeout 120 python bench.py --workload multihot --no-cpu-baseline > gpurun_out/r2_b38_multihot.json 2> gpurun_out/r2_b38_multihot.err; echo "bench rc=$?"
timeout 120 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b38_wide.json 2> gpurun_out/r2_b38_wide.err; echo "bench rc=$?"
WD_STEP_TRACE=1 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/r2_b38_trace.json 2> gpurun_out/r2_b38_trace.err
grep -h "timeline" gpurun_out/r2_b38_trace.err | tail -1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/r2_launches_v3.csv python bench.py --steps 20 --warmup 30 --no-cpu-baseline > gpurun_out/r2_ncu_c.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches_v3.csv | tail -3
python - <<'PY'
import json
for f in ("r2_b38","r2_b38_multihot","r2_b38_wide"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["launches_per_step"], d.get("roofline",{}).get("frac"), d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
