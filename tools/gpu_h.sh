#!/bin/bash
# call H: after de-inlining the shared helpers / rolling the pixel loop — parity, workloads, launch lists, stall profiles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_multi_gpu.py::test_direct_sharded_contexts_share_one_framebuffer > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q -k direct_sharded > gpurun_out/h_pytest_mgpu.log 2>&1; tail -3 gpurun_out/h_pytest_mgpu.log
for fm in 0 32; do
  WRCU_FLAT_MAX=$fm timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_flat$fm.json 2> gpurun_out/h_bench_flat$fm.err; echo "bench flat_max=$fm rc=$?"
done
for w in composite clip_rects text video_nv12 b_prime images gradients box_shadow blur page; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/h_workloads.jsonl 2>> gpurun_out/h_workloads.err
done
cat gpurun_out/h_workloads.jsonl | cut -c1-420
for w in composite clip_rects page; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu_$w.log 2>&1
done
SECS="--section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats"
timeout 300 ncu $SECS --clock-control none -k regex:wr_composite_copy -s 3 -c 1 -o gpurun_out/h_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu_copy.log 2>&1
timeout 300 ncu $SECS --clock-control none -k regex:"wr_raster<ClipRectShader" -s 6 -c 1 -o gpurun_out/h_prof_cliprect python bench.py --workload clip_rects --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu_cliprect.log 2>&1
timeout 300 ncu $SECS --clock-control none -k regex:wr_setup_clip_rectangle -s 6 -c 1 -o gpurun_out/h_prof_setup python bench.py --workload clip_rects --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu_setup.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
