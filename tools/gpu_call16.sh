#!/bin/bash
# round-2 GPU call 16: split-K factor chosen by the makespan model (A/B: CRNN_KSPLIT=old); host profile of the e2e step
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_shapes.py -m gpu -q --timeout=600 > gpurun_out/c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
tail -6 gpurun_out/c16_pytest.log
timeout 400 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c16_bench_new.json 2> gpurun_out/c16_bench_new.err
CRNN_KSPLIT=old timeout 400 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c16_bench_oldksplit.json 2> gpurun_out/c16_bench_oldksplit.err
python - <<'PY'
import json
for f in ("new", "oldksplit"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c16_bench_{f}.json") if l.startswith("{")][-1])
        t = d.get("train_step") or {}
        s = t.get("stages_ms") or {}
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "train", t.get("ms_per_step"),
              {k: v for k, v in s.items() if "wgrad" in k or k in ("conv5_bwd", "zero+logits_bwd")})
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 python tools/e2e_profile.py 30 > gpurun_out/c16_e2e_profile.txt 2>&1
head -60 gpurun_out/c16_e2e_profile.txt
