#!/bin/bash
# call I: deferred submission (one set-up launch per flush), A/B against WRCU_IMMEDIATE=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/i_pytest.log
tail -5 gpurun_out/i_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/i_bench.err
for imm in 0 1; do
for w in composite clip_rects text video_nv12 b_prime images gradients box_shadow blur page; do
  WRCU_IMMEDIATE=$imm timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/i_workloads_imm$imm.jsonl 2>> gpurun_out/i_workloads.err
done
echo "== immediate=$imm"; python - <<PY
import json
for l in open("gpurun_out/i_workloads_imm$imm.jsonl"):
    d=json.loads(l); print("%-12s flushed %.3f ms  warm %.3f  pipelined %.3f  launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
done
for w in clip_rects page; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/i_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/i_ncu_$w.log 2>&1
done
SECS="--section SpeedOfLight --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy --section SchedulerStats"
timeout 300 ncu $SECS --clock-control none -k regex:wr_raster -s 3 -c 1 -o gpurun_out/i_prof_cliprect python bench.py --workload clip_rects --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/i_ncu_cliprect.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
