#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/d_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/d_bench.err
for w in composite clip_rects text images; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/d_workloads.jsonl 2>> gpurun_out/d_workloads.err
done
cat gpurun_out/d_workloads.jsonl
for w in composite images; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/d_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/d_ncu_$w.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wr_raster<CompositeShader" -s 8 -c 4 -o gpurun_out/d_prof_generic python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/d_ncu_generic.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wr_composite_copy<0>" -s 3 -c 1 -o gpurun_out/d_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/d_ncu_copy.log 2>&1
echo done
