#!/bin/bash
mkdir -p gpurun_out
for i in 0 7 8 9 1 2 3 4 5 6; do timeout 60 tools/probe/tma_probe $i 2>&1 | grep -v "^probe:" ; done > gpurun_out/c_probe.log 2>&1; cat gpurun_out/c_probe.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c_pytest.log
timeout 600 python -m pytest tests -m gpu -q -k "behind_opaque or occluded or occluders or composite" > gpurun_out/c_pytest_focus.log 2>&1; echo "focus rc=$?"; tail -12 gpurun_out/c_pytest_focus.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
for w in composite clip_rects text images; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/c_workloads.jsonl 2>> gpurun_out/c_workloads.err
done
cat gpurun_out/c_workloads.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c_launches_composite.csv python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_composite.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wr_composite_copy -s 4 -c 2 -o gpurun_out/c_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_copy.log 2>&1
echo done
