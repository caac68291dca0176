#!/bin/bash
# run P: fill class + ragged-path MLP + clears ahead of the set-up launch; per-line profiles of the set-up, text and YUV kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/p_pytest.log; tail -3 gpurun_out/p_pytest.log | cut -c1-200
for w in composite clip_rects text video_nv12 gradients box_shadow page; do
  timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/p_workloads.jsonl 2>> gpurun_out/p_workloads.err
done
WRCU_GLYPH_MAJOR=0 timeout 200 python bench.py --workload text --steps 10 --no-cpu-baseline > gpurun_out/p_text_tileonly.json 2>> gpurun_out/p_workloads.err
WRCU_GLYPH_MAJOR=0 timeout 200 python bench.py --workload page --steps 10 --no-cpu-baseline > gpurun_out/p_page_tileonly.json 2>> gpurun_out/p_workloads.err
python - <<PY
import json
for l in open("gpurun_out/p_workloads.jsonl"):
    d=json.loads(l)
    print("%-12s %.3f ms  (warm %.3f, pipelined %.3f) launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
for f in ("p_text_tileonly","p_page_tileonly"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f, "%.3f ms"%d["ms_per_step"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/p_launches_composite.csv python bench.py --workload composite --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_composite.log 2>&1
prof() {  # name workload kernel-regex skip
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$3 -s $4 -c 1 -o /tmp/p_prof_$1 python bench.py --workload $2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_$1.log 2>&1
  ncu -i /tmp/p_prof_$1.ncu-rep --page raw --csv > gpurun_out/p_prof_$1.raw.csv 2>/dev/null
  ncu -i /tmp/p_prof_$1.ncu-rep --page source --print-source cuda --csv 2>/dev/null | cut -d, -f1-9 > gpurun_out/p_prof_$1.cuda.csv
}
prof setup_composite composite wr_setup_multi 1
prof setup_text text wr_setup_multi 1
prof text text TextShader 1
prof glyphs text wr_raster_glyphs 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/p_launches_text.csv python bench.py --workload text --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_ncu_text_l.log 2>&1
prof yuv video_nv12 CompositeYuvShader 1
prof copy0 composite "wr_composite_copy<\(int\)0" 1
du -sh gpurun_out
echo done
