#!/bin/bash
mkdir -p gpurun_out
for env in "" "WRCU_GLYPH_MAJOR=0" "WRCU_STREAMS=1" "WRCU_PDL=0" "WRCU_EARLY_CLEAR=0" "WRCU_IMMEDIATE=1"; do
  echo "=== $env" | tee -a gpurun_out/r_debug.log
  env $env timeout 200 python tools/debug_update_path.py 2>&1 | tail -22 | tee -a gpurun_out/r_debug.log
done
for i in 1 2 3; do timeout 200 python bench.py --workload page --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('page', round(d['ms_per_step'],3))"; done
