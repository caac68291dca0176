#!/bin/bash
# multi-GPU call: N = $1 ranks on one node (gpurun --gpus N): config-B weak scaling + config E (one 8K frame sharded by tile)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/m${N}_smi.txt
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/m${N}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/m${N}_bench.json") if l.startswith("{")][-1])
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("config_e"), indent=1))
PY
echo done
