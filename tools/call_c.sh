set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_bundled.py -q -k "crelu or wide_only or multihot or column_ids or bundled or run_to_run" > gpurun_out/r2_t33.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/r2_t33.log
timeout 200 python bench.py --workload multihot --no-cpu-baseline > gpurun_out/r2_b33_multihot.json 2> gpurun_out/r2_b33_multihot.err; echo "bench rc=$?"
timeout 200 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b33_wide.json 2> gpurun_out/r2_b33_wide.err; echo "bench rc=$?"
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_b33.json 2> gpurun_out/r2_b33.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_b33","r2_b33_multihot","r2_b33_wide"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d.get("roofline",{}).get("frac"))
        print({k: v for k, v in d["kernels"]["phases_ms"].items() if v > 0.03})
    except Exception as e: print(f, "ERR", e)
PY
