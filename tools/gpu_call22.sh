#!/bin/bash
# round-2 GPU call 22: last check of the committed build -- full GPU suite + the default bench line
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1300 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c22_pytest.log
tail -4 gpurun_out/c22_pytest.log
timeout 500 python bench.py > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
echo "bench rc=$?" >> gpurun_out/c22_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c22_bench.json") if l.startswith("{")][-1])
    t = d.get("train_step") or {}
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k: v["value"] for k, v in d["e2e"]["variants"].items()}, "train", t.get("ms_per_step"),
          "delta", d["ctc_loss_delta"]["rel"], "decode", d["decode_equality"].get("agreement_unfiltered"), d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
