#!/bin/bash
# call K: side streams per render target + staged depth prepass; full parity, A/B over WRCU_STREAMS, config E on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/k_pytest.log
tail -6 gpurun_out/k_pytest.log | cut -c1-300
for ns in 8 1 16; do
for w in composite clip_rects text images blur page; do
  WRCU_STREAMS=$ns timeout 120 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/k_workloads_ns$ns.jsonl 2>> gpurun_out/k_workloads.err
done
echo "== streams=$ns"; python - <<PY
import json
for l in open("gpurun_out/k_workloads_ns$ns.jsonl"):
    d=json.loads(l); print("%-12s flushed %.3f ms  warm %.3f  pipelined %.3f  launches %d"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"]))
PY
done
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/k_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/k_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
print([(e["layers"], round(e["raster_kernel_ms"]*1000,1), round(e.get("dram_frac_of_hbm",0),3)) for e in d["roofline_sweep"]])
print(d.get("config_e"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/k_launches_page.csv python bench.py --workload page --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu_page.log 2>&1
du -sh gpurun_out
echo done
