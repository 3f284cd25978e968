set -x
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 300 --csv --log-file gpurun_out/r2_launches_multihot.csv python bench.py --workload multihot --steps 20 --warmup 30 --no-cpu-baseline > gpurun_out/r2_ncu_mh.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches_multihot.csv | tail -45
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 150 --csv --log-file gpurun_out/r2_launches_wide.csv python bench.py --workload wide --steps 20 --warmup 30 --no-cpu-baseline > gpurun_out/r2_ncu_w.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches_wide.csv | tail -30
