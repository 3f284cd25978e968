set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -k "wide_only or run_to_run or criteo or multihot or dense_exchange or optimizers" > gpurun_out/r2_t37.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r2_t37.log
timeout 200 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b37_wide.json 2> gpurun_out/r2_b37_wide.err; echo "bench rc=$?"
WD_SORT_NO_LOCAL=1 timeout 200 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b37_wide_nolocal.json 2> gpurun_out/r2_b37_wide_nolocal.err; echo "bench rc=$?"
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_b37.json 2> gpurun_out/r2_b37.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_b37","r2_b37_wide","r2_b37_wide_nolocal"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d.get("roofline",{}).get("frac"))
        print({k: v for k, v in d["kernels"]["phases_ms"].items() if v > 0.03})
    except Exception as e: print(f, "ERR", e)
PY
