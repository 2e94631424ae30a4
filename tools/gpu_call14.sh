#!/bin/bash
# round-2 GPU call 14: conv1 weight gradient on the tensor cores, conv3_1 ReLU backward + conv4_1 BN-backward sums fused into the
# producing data-gradient epilogues.  Gradient parity first, then A/B of each switch inside one box.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_shapes.py -m gpu -q --timeout=600 -x > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
tail -15 gpurun_out/c14_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c14_bench_$name.json 2> gpurun_out/c14_bench_$name.err
}
run new CRNN_NOP=1
run old_conv1wg CRNN_CONV1_WGRAD=simt
run old_relu CRNN_RELU_FUSE=0
run old_bn CRNN_BN_FUSE=0
python - <<'PY'
import json
for f in ("new", "old_conv1wg", "old_relu", "old_bn"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c14_bench_{f}.json") if l.startswith("{")][-1])
        s = d["train_step"]["stages_ms"]
        print(f, d["value"], d["ms_per_step"], "train", d["train_step"]["ms_per_step"],
              {k: s[k] for k in ("conv4_2_dgrad", "bn4_1_bwd", "conv3_2_dgrad", "conv3_1_bwd_elem", "conv1_wgrad")})
    except Exception as e:
        print(f, "failed", e)
PY
