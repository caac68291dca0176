#!/bin/bash
# call J: PDL chain + copy-kernel split; kernel durations of the deferred path (CSV only: reports stay on the box)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | grep -E "Error|passed|failed|assert" | head -12 > gpurun_out/j_mgpu.log; cat gpurun_out/j_mgpu.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_multi_gpu.py::test_direct_sharded_contexts_share_one_framebuffer > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log
for pdl in 1 0; do
for w in composite clip_rects text video_nv12 b_prime images gradients box_shadow blur page; do
  WRCU_PDL=$pdl timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/j_workloads_pdl$pdl.jsonl 2>> gpurun_out/j_workloads.err
done
echo "== pdl=$pdl"; python - <<PY
import json
for l in open("gpurun_out/j_workloads_pdl$pdl.jsonl"):
    d=json.loads(l); print("%-12s flushed %.3f ms  warm %.3f  pipelined %.3f  launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
done
for w in gradients clip_rects composite page; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_$w.log 2>&1
done
SECS="--section SpeedOfLight --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats"
timeout 300 ncu $SECS --clock-control none -k regex:wr_setup_multi -s 3 -c 1 -o /tmp/j_prof_setup python bench.py --workload gradients --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_setup.log 2>&1
ncu -i /tmp/j_prof_setup.ncu-rep --page raw --csv > gpurun_out/j_prof_setup_gradients.raw.csv 2>/dev/null
timeout 300 ncu $SECS --section SourceCounters --clock-control none -k regex:wr_raster -s 3 -c 1 -o /tmp/j_prof_clip python bench.py --workload clip_rects --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_clip.log 2>&1
ncu -i /tmp/j_prof_clip.ncu-rep --page raw --csv > gpurun_out/j_prof_cliprect.raw.csv 2>/dev/null
ncu -i /tmp/j_prof_clip.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-12 | head -4000 > gpurun_out/j_prof_cliprect.source.csv
timeout 300 ncu $SECS --clock-control none -k regex:wr_composite_copy -s 3 -c 1 -o /tmp/j_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_copy.log 2>&1
ncu -i /tmp/j_prof_copy.ncu-rep --page raw --csv > gpurun_out/j_prof_copy.raw.csv 2>/dev/null
du -sh gpurun_out
echo done
