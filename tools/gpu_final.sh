#!/bin/bash
# round 2 final evidence run: parity, smoke, the bench line, the reference arm (all / 8 / 1 cores), every workload with the
# SWGL baseline, launch lists, ncu captures exported as CSV (reports stay on the box)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/z_smi.txt 2>&1; nproc >> gpurun_out/z_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/z_pytest.log; tail -3 gpurun_out/z_pytest.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log
timeout 600 python bench.py --config-e > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/z_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_ref_all.json 2> gpurun_out/z_ref.err; echo "ref all rc=$?"
timeout 400 python bench.py --impl reference --ref-cores 8 --steps 3 --warmup 1 > gpurun_out/z_ref_8.json 2>> gpurun_out/z_ref.err
timeout 400 python bench.py --impl reference --ref-cores 1 --steps 2 --warmup 1 --ref-rects 100 > gpurun_out/z_ref_1.json 2>> gpurun_out/z_ref.err
for w in page composite clip_rects text video_nv12 gradients box_shadow images blur b_prime; do
  timeout 200 python bench.py --workload $w --steps 10 >> gpurun_out/z_workloads.jsonl 2>> gpurun_out/z_workloads.err
done
for g in 4 6 8; do WRCU_GLYPH_CTAS=$g timeout 200 python bench.py --workload text --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('text glyph_ctas=$g', round(d['ms_per_step'],3))"; done
for w in video_nv12 gradients; do WRCU_STRIP=0 timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline >> gpurun_out/z_workloads_strip0.jsonl 2>/dev/null; done
python - <<PY
import json
for f in ("z_ref_all","z_ref_8","z_ref_1"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f, round(d["value"],1), "Mpix/s cores", d["config"]["cores"])
    except Exception as e: print(f, "ERR", e)
for l in open("gpurun_out/z_workloads.jsonl"):
    d=json.loads(l); cb=d.get("cpu_baseline",{})
    print("%-12s %.3f ms  (warm %.3f, pipelined %.3f) launches %d | SWGL 1 core %.2f ms"%(d["config"]["workload"], d["ms_per_step"], d["ms_warm_l2"], d["ms_pipelined"], d["gpu_launches"], 1e3/cb["value"] if cb else -1))
PY
for w in composite page text; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/z_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/z_ncu_$w.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/z_launches_configB.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep > gpurun_out/z_ncu_configB.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:wr_raster_solid_premult -s 2 -c 1 -o /tmp/z_prof_premult python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep > gpurun_out/z_ncu_premult.log 2>&1
ncu -i /tmp/z_prof_premult.ncu-rep --page raw --csv > gpurun_out/ncu_full_solid_premult_r02.raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:wr_raster_solid --csv --log-file gpurun_out/z_sweep_dram.csv python bench.py --sweep-only --steps 2 > gpurun_out/z_ncu_sweep.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:wr_composite_copy -s 3 -c 2 -o /tmp/z_prof_copy python bench.py --workload composite --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/z_ncu_copy.log 2>&1
ncu -i /tmp/z_prof_copy.ncu-rep --page raw --csv > gpurun_out/ncu_full_composite_copy_r02.raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:wr_raster_solid_flat -s 2 -c 1 -o /tmp/z_prof_flat python bench.py --sweep-only --steps 2 > gpurun_out/z_ncu_flat.log 2>&1
ncu -i /tmp/z_prof_flat.ncu-rep --page raw --csv > gpurun_out/ncu_full_solid_flat_r02.raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:wr_raster_glyphs -s 1 -c 1 -o /tmp/z_prof_glyphs python bench.py --workload text --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/z_ncu_glyphs.log 2>&1
ncu -i /tmp/z_prof_glyphs.ncu-rep --page raw --csv > gpurun_out/ncu_full_text_glyphs_r02.raw.csv 2>/dev/null
du -sh gpurun_out
echo done
