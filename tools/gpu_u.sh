#!/bin/bash
# run U: strip mode of the tile kernel (row state kept across the tiles of a row) for video / gradients, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "yuv or gradient or composite or golden or page or video or host_renderer or gl_shim" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/u_pytest.log; tail -3 gpurun_out/u_pytest.log | cut -c1-200
grep -E "^FAILED" gpurun_out/u_pytest.log | head -20
for w in video_nv12 gradients; do
  for i in 1 2; do
  timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/u_workloads.jsonl 2>> gpurun_out/u_workloads.err
  WRCU_STRIP=0 timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/u_workloads_strip0.jsonl 2>> gpurun_out/u_workloads.err
  done
done
python - <<PY
import json
for f in ("u_workloads","u_workloads_strip0"):
    print(f)
    for l in open("gpurun_out/%s.jsonl"%f):
        try: d=json.loads(l)
        except Exception: continue
        print("  %-12s %.3f ms  (warm %.3f, pipelined %.3f) launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
for w in video_nv12 gradients; do
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'^wr_raster$' -c 6 --csv --log-file gpurun_out/u_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/u_ncu_$w.log 2>&1
WRCU_STRIP=0 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'^wr_raster$' -c 6 --csv --log-file gpurun_out/u_launches_${w}_strip0.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/u_ncu_${w}0.log 2>&1
done
grep -h "wr_raster" gpurun_out/u_launches_*.csv | awk -F'","' '{print FILENAME, $5, $(NF-2), $NF}' | head -40
echo done
