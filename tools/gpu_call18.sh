#!/bin/bash
# round-2 GPU call 18 (gpurun --gpus 2): data-parallel equivalence tests and the default 2-GPU bench line on the final build
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout=500 > gpurun_out/c18_pytest_dp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c18_pytest_dp.log
tail -5 gpurun_out/c18_pytest_dp.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 500)) \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c18_bench_${N}gpu.json 2> gpurun_out/c18_bench_${N}gpu.err
echo "rc=$?" >> gpurun_out/c18_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c18_bench_${N}gpu.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "train", d.get("train_step", {}).get("ms_per_step"), d.get("global_batch_bn_forward"))
except Exception as e:
    print("failed", e)
PY
tail -3 gpurun_out/c18_bench_${N}gpu.err
