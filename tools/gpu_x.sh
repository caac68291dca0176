#!/bin/bash
# sanity of the rebuilt library: smoke, goldens + shim on the GPU, the default bench line
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q -k "golden or gl_shim or text or yuv" 2>&1 | tail -2 | cut -c1-200
timeout 600 python bench.py > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/x_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["traffic"], d["roofline"].get("traffic_source"), d["workloads"])
PY
