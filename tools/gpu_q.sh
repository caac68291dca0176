#!/bin/bash
# run Q: glyph-major text with resolve flags (no second pass unless a glyph cannot be drawn), plain launches after event
# waits, WRCU_EARLY_CLEAR A/B; per-line profiles (CUDA source page kept whole)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log | cut -c1-200
grep -E "^FAILED" gpurun_out/q_pytest.log | head -20
for w in composite clip_rects text video_nv12 gradients page; do
  timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/q_workloads.jsonl 2>> gpurun_out/q_workloads.err
done
for w in text page gradients clip_rects; do
  WRCU_EARLY_CLEAR=0 timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/q_workloads_ec0.jsonl 2>> gpurun_out/q_workloads.err
done
WRCU_GLYPH_MAJOR=0 timeout 200 python bench.py --workload page --steps 10 --no-cpu-baseline >> gpurun_out/q_workloads_gm0.jsonl 2>> gpurun_out/q_workloads.err
python - <<PY
import json
for f in ("q_workloads","q_workloads_ec0","q_workloads_gm0"):
    print(f)
    for l in open("gpurun_out/%s.jsonl"%f):
        try: d=json.loads(l)
        except Exception: continue
        print("  %-12s %.3f ms  (warm %.3f, pipelined %.3f) launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
for w in text page; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/q_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/q_ncu_$w.log 2>&1
done
prof() {  # name workload kernel-regex skip
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$3 -s $4 -c 1 -o /tmp/q_prof_$1 python bench.py --workload $2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/q_ncu_$1.log 2>&1
  ncu -i /tmp/q_prof_$1.ncu-rep --page raw --csv > gpurun_out/q_prof_$1.raw.csv 2>/dev/null
  ncu -i /tmp/q_prof_$1.ncu-rep --page source --print-source cuda --csv 2>/dev/null | gzip -9 > gpurun_out/q_prof_$1.source.csv.gz
}
prof glyphs text wr_raster_glyphs 1
prof setup_text text wr_setup_multi 1
prof yuv video_nv12 '^wr_raster$' 1
prof clip clip_rects '^wr_raster$' 1
du -sh gpurun_out
echo done
