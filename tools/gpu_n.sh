#!/bin/bash
# call N: row tables for short commands (text), side-grid rule; full parity + workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/n_pytest.log | cut -c1-200
for w in page composite clip_rects text images video_nv12 gradients box_shadow blur b_prime; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/n_workloads.jsonl 2>> gpurun_out/n_workloads.err
done
python - <<PY
import json
for l in open("gpurun_out/n_workloads.jsonl"):
    d=json.loads(l); print("%-12s flushed %.3f ms  warm %.3f  pipelined %.3f  launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/n_launches_text.csv python bench.py --workload text --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/n_ncu_text.log 2>&1
grep -E "wr_raster|wr_setup" gpurun_out/n_launches_text.csv | tail -4 | cut -d, -f5,15- | cut -c1-200
echo done
