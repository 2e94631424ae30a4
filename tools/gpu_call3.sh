#!/bin/bash
# round-2 GPU call 3 (gpurun --gpus N): data-parallel correctness (N = 2) and the scaling variants
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c3_topo.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q --timeout=300 -k "ctc or gradients" > gpurun_out/c3_pytest_ctc.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/c3_pytest_ctc.log
  tail -4 gpurun_out/c3_pytest_ctc.log
  timeout 100 python tools/ctc_bench.py > gpurun_out/c3_ctc.log 2>&1; cat gpurun_out/c3_ctc.log
  timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout=500 > gpurun_out/c3_pytest_dp.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/c3_pytest_dp.log
  tail -5 gpurun_out/c3_pytest_dp.log
fi
run() { # name, extra flags
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 500)) \
      bench.py --gpus $N --steps 20 --warmup 5 $2 > gpurun_out/c3_bench_${N}gpu_$1.json 2> gpurun_out/c3_bench_${N}gpu_$1.err
  echo "rc=$?" >> gpurun_out/c3_bench_${N}gpu_$1.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c3_bench_${N}gpu_$1.json") if l.startswith("{")][-1])
    print("$1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "train", d.get("train_step", {}).get("ms_per_step"))
except Exception as e:
    print("$1 failed", e)
PY
}
run default ""
run sync_bn_forward "--sync-bn-forward"
if [ "$N" = "2" ]; then
  run overlap "--overlap"
  run nccl_bn "--sync-bn-forward --no-peer-memory"
fi
tail -3 gpurun_out/c3_bench_${N}gpu_default.err
