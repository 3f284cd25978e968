set -x
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_bundled.py -q > gpurun_out/r2_t35.log 2>&1; echo "pytest rc=$?"
tail -n 8 gpurun_out/r2_t35.log
timeout 200 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b35_wide.json 2> gpurun_out/r2_b35_wide.err; echo "bench rc=$?"
WD_SORT_DIGIT_BITS=8 timeout 200 python bench.py --workload wide --no-cpu-baseline > gpurun_out/r2_b35_wide_d8.json 2> gpurun_out/r2_b35_wide_d8.err; echo "bench rc=$?"
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_b35.json 2> gpurun_out/r2_b35.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_b35","r2_b35_wide","r2_b35_wide_d8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d.get("roofline",{}).get("frac"))
        print({k: v for k, v in d["kernels"]["phases_ms"].items() if v > 0.03})
    except Exception as e: print(f, "ERR", e)
PY
