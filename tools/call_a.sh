set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_bundled.py -x -q > gpurun_out/r2_t31.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r2_t31.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_b31.json 2> gpurun_out/r2_b31.err; echo "bench rc=$?"
WD_NO_FUSED_ROW_APPLY=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_b31_nofuse.json 2> gpurun_out/r2_b31_nofuse.err
WD_STEP_TRACE=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_b31_trace.json 2> gpurun_out/r2_b31_trace.err
python - <<'PY'
import json
for f in ("r2_b31","r2_b31_nofuse","r2_b31_trace"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
grep -h "timeline" gpurun_out/r2_b31_trace.err | tail -1
