#!/bin/bash
# round-2 GPU call 9: BN backward with recomputed dy, conv2 weight gradient with swapped operands, 128-bit reds in the wgrad epilogues
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
timeout 400 python bench.py --no-decode-eq > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
echo "bench rc=$?" >> gpurun_out/c9_bench.err
CRNN_CONV2_WGRAD=old timeout 400 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c9_bench_oldwgrad.json 2> gpurun_out/c9_bench_oldwgrad.err
tail -6 gpurun_out/c9_pytest.log
python - <<'PY'
import json
for f in ("c9_bench", "c9_bench_oldwgrad"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], "train", d["train_step"]["ms_per_step"], d["train_step"]["stages_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
