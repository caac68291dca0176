#!/bin/bash
# run W: copy kernels — ring depth x resident CTAs (WRCU_COPY_STAGES / WRCU_COPY_CTAS, WRCU_BLEND_STAGES / WRCU_BLEND_CTAS)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "composite or page or gl_shim or host_renderer" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/w_pytest.log | cut -c1-200
run() {  # label env...
  local label=$1; shift
  local ms=$(env "$@" timeout 200 python bench.py --workload composite --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_step'],4), round(d['ms_pipelined'],4))")
  env "$@" timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wr_composite_copy -s 8 -c 4 --csv --log-file gpurun_out/w_copy_$label.csv python bench.py --workload composite --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
  local ks=$(grep wr_composite_copy gpurun_out/w_copy_$label.csv | awk -F'","' '{gsub(/"/,"",$NF); printf "%s:%s ", substr($5,1,21), $NF}')
  echo "$label frame(ms, pipelined) $ms | $ks" | tee -a gpurun_out/w_results.txt
}
run st4c3 WRCU_COPY_STAGES=4 WRCU_COPY_CTAS=3
run st6c2 WRCU_COPY_STAGES=6 WRCU_COPY_CTAS=2
run st8c1 WRCU_COPY_STAGES=8 WRCU_COPY_CTAS=1
run st5c2 WRCU_COPY_STAGES=5 WRCU_COPY_CTAS=2
run st3c4 WRCU_COPY_STAGES=3 WRCU_COPY_CTAS=4
run st4c2 WRCU_COPY_STAGES=4 WRCU_COPY_CTAS=2
run st4c1 WRCU_COPY_STAGES=4 WRCU_COPY_CTAS=1
run b4c1 WRCU_BLEND_STAGES=4 WRCU_BLEND_CTAS=1
run b2c3 WRCU_BLEND_STAGES=2 WRCU_BLEND_CTAS=3
run b3c1 WRCU_BLEND_STAGES=3 WRCU_BLEND_CTAS=1
run b4c2 WRCU_BLEND_STAGES=4 WRCU_BLEND_CTAS=2
echo done
