#!/bin/bash
# what the driver's scaling run does: bench.py at N GPUs with default flags (one JSON line on rank 0)
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 500)) \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_${N}gpu.json 2> gpurun_out/scale_${N}gpu.err
echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/scale_${N}gpu.json") if l.startswith("{")][-1])
print("$N GPUs", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sync-bn fwd", d.get("global_batch_bn_forward", {}).get("ms_per_step"), "train", d["train_step"]["ms_per_step"], d["train_step"]["images_per_s"])
PY
