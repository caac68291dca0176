#!/bin/bash
# round 2, session 2, call G: parity suite (perspective, split composite, copy kernel rewrite), flat/tile crossover,
# warm vs flushed workload timings, ncu full of the small-batch kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/g_pytest.log
tail -8 gpurun_out/g_pytest.log
for fm in 0 32; do
  WRCU_FLAT_MAX=$fm timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_flat$fm.json 2> gpurun_out/g_bench_flat$fm.err; echo "bench flat_max=$fm rc=$?"
done
for w in composite clip_rects text video_nv12 b_prime images gradients box_shadow blur; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/g_workloads.jsonl 2>> gpurun_out/g_workloads.err
done
cat gpurun_out/g_workloads.jsonl
for w in composite clip_rects; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/g_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_$w.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wr_composite_copy -s 3 -c 2 -o gpurun_out/g_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_copy.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wr_raster<ClipRectShader" -s 6 -c 2 -o gpurun_out/g_prof_cliprect python bench.py --workload clip_rects --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_cliprect.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wr_setup_clip_rectangle -s 6 -c 2 -o gpurun_out/g_prof_setup python bench.py --workload clip_rects --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_ncu_setup.log 2>&1
echo done
