#!/bin/bash
# one 8-GPU box: config-B weak scaling + config E strong scaling at N = 8, 4, 2, 1
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/m_smi.txt
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
for N in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err; echo "N=$N bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/m${N}_bench.json") if l.startswith("{")][-1])
e=d.get("config_e") or {}
print("N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "| config E ms", e.get("ms_per_frame"), "1gpu", e.get("one_gpu_ms_per_frame"), "speedup", e.get("speedup_vs_one_gpu"), "crc ok", e.get("matches_single_gpu"), e.get("error"))
PY
done
echo done
