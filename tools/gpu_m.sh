#!/bin/bash
# call M: side-stream grid sizing A/B, new reference pins on the CUDA tier, default bench line with the workloads summary
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_golden.py tests/test_multi_gpu.py tests/test_host_renderer.py -m gpu -q > gpurun_out/mm_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/mm_pytest.log | cut -c1-200
for sc in 1 2 3; do
for w in page composite clip_rects images; do
  WRCU_SIDE_CTAS=$sc timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/mm_workloads_sc$sc.jsonl 2>> gpurun_out/mm_workloads.err
done
echo "== side ctas/SM=$sc"; python - <<PY
import json
for l in open("gpurun_out/mm_workloads_sc$sc.jsonl"):
    d=json.loads(l); print("%-12s flushed %.3f ms  warm %.3f  pipelined %.3f  launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
done
timeout 400 python bench.py --steps 10 --warmup 3 --config-e > gpurun_out/mm_bench.json 2> gpurun_out/mm_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/mm_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/mm_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"))
print(json.dumps(d.get("workloads")))
print((d.get("config_e") or {}).get("ms_per_frame"))
PY
echo done
