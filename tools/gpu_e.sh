#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/e_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/e_bench.err
for w in composite clip_rects text images; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/e_workloads.jsonl 2>> gpurun_out/e_workloads.err
done
cat gpurun_out/e_workloads.jsonl
for w in composite images clip_rects; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/e_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_$w.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k wr_composite_copy -s 9 -c 3 -o gpurun_out/e_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_copy.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k wr_raster -s 12 -c 4 -o gpurun_out/e_prof_generic python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_generic.log 2>&1
echo done
