#!/bin/bash
# run S: glyph-major with per-lane row walks (rows under 4 pixels), wide YUV variant A/B; source pages kept whole (gz)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s_pytest.log; tail -3 gpurun_out/s_pytest.log | cut -c1-200
grep -E "^FAILED" gpurun_out/s_pytest.log | head -20
for w in text page video_nv12 clip_rects composite; do
  timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/s_workloads.jsonl 2>> gpurun_out/s_workloads.err
done
WRCU_YUV_WIDE=1 timeout 200 python bench.py --workload video_nv12 --steps 10 --no-cpu-baseline >> gpurun_out/s_workloads_wide.jsonl 2>> gpurun_out/s_workloads.err
python - <<PY
import json
for f in ("s_workloads","s_workloads_wide"):
    print(f)
    for l in open("gpurun_out/%s.jsonl"%f):
        try: d=json.loads(l)
        except Exception: continue
        print("  %-12s %.3f ms  (warm %.3f, pipelined %.3f) launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s_launches_text.csv python bench.py --workload text --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s_ncu_text.log 2>&1
prof() {  # name workload kernel-regex skip
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$3 -s $4 -c 1 -o /tmp/s_prof_$1 python bench.py --workload $2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s_ncu_$1.log 2>&1
  ncu -i /tmp/s_prof_$1.ncu-rep --page raw --csv > gpurun_out/s_prof_$1.raw.csv 2>/dev/null
  ncu -i /tmp/s_prof_$1.ncu-rep --page source --print-source cuda --csv 2>/dev/null | gzip -9 > gpurun_out/s_prof_$1.cuda.csv.gz
}
prof glyphs text wr_raster_glyphs 1
prof yuv video_nv12 '^wr_raster$' 1
prof setup_text text wr_setup_multi 1
du -sh gpurun_out
echo done
