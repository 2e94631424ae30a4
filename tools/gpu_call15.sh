#!/bin/bash
# round-2 GPU call 15: conv1 weight-gradient kernel with loads issued two tiles ahead; Session device prefetch (next ring slot copied
# host->device during the current step) -- test + e2e A/B
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_shapes.py -m gpu -q --timeout=600 -x > gpurun_out/c15_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c15_pytest.log
tail -8 gpurun_out/c15_pytest.log
timeout 400 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c15_bench_new.json 2> gpurun_out/c15_bench_new.err
CRNN_BENCH_NO_DEVICE_PREFETCH=1 timeout 400 python bench.py --no-decode-eq --no-cpu-baseline --no-train > gpurun_out/c15_bench_noprefetch.json 2> gpurun_out/c15_bench_noprefetch.err
python - <<'PY'
import json
for f in ("new", "noprefetch"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c15_bench_{f}.json") if l.startswith("{")][-1])
        t = d.get("train_step") or {}
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["feed"][-90:], {k: v["value"] for k, v in d["e2e"]["variants"].items()},
              "train", t.get("ms_per_step"), (t.get("stages_ms") or {}).get("conv1_wgrad"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/c15_bench_new.err
