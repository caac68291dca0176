#!/bin/bash
# round 2, GPU call A: parity suite, bench with the layer sweep, the bandwidth-bound workloads, launch lists
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
for w in composite clip_rects text video_nv12 b_prime images; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/a_workloads.jsonl 2>> gpurun_out/a_workloads.err
done
cat gpurun_out/a_workloads.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a_launches_composite.csv python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_composite.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wr_composite_copy -c 2 -o gpurun_out/a_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_copy.log 2>&1
echo done
