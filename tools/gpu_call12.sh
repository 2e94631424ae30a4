#!/bin/bash
# round-2 GPU call 12: conv3_1 data gradient through the swapped-operand kernel as well (A/B: CRNN_CONV2_DGRAD=old)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
timeout 400 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "bench rc=$?" >> gpurun_out/c12_bench.err
CRNN_CONV2_DGRAD=old timeout 400 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c12_bench_olddgrad.json 2> gpurun_out/c12_bench_olddgrad.err
tail -6 gpurun_out/c12_pytest.log
python - <<'PY'
import json
for f in ("c12_bench", "c12_bench_olddgrad"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], "train", d["train_step"]["ms_per_step"], d["train_step"]["stages_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
