#!/bin/bash
# round 2, session 2, call F: full parity suite, the bench line (sweep + config E), the other workloads,
# launch lists and full ncu captures of the copy-class composite kernel and the shallow solid kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/f_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout 500 python bench.py --steps 10 --warmup 3 --config-e > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err
for w in composite clip_rects text video_nv12 b_prime images gradients box_shadow blur; do
  timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/f_workloads.jsonl 2>> gpurun_out/f_workloads.err
done
cat gpurun_out/f_workloads.jsonl
for w in composite clip_rects; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/f_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_$w.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_launches_configB.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep > gpurun_out/f_ncu_configB.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wr_composite_copy -s 3 -c 2 -o gpurun_out/f_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_copy.log 2>&1
echo done
