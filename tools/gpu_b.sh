#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/probe/tma_probe > gpurun_out/b_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/b_probe.log
timeout 300 python -m pytest tests -m gpu -x -q -k "composite" > gpurun_out/b_pytest_composite.log 2>&1; echo "pytest composite rc=$?"; tail -4 gpurun_out/b_pytest_composite.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/b_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
for w in composite clip_rects; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/b_workloads.jsonl 2>> gpurun_out/b_workloads.err
done
cat gpurun_out/b_workloads.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/b_launches_composite.csv python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_composite.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wr_composite_copy -s 4 -c 2 -o gpurun_out/b_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_copy.log 2>&1
echo done
