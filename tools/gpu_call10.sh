#!/bin/bash
# round-2 GPU call 10: conv2 data gradient with swapped operands (A/B against the position-major N = 64 kernel)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 400 python bench.py > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
echo "bench rc=$?" >> gpurun_out/c10_bench.err
CRNN_CONV2_DGRAD=old timeout 400 python bench.py --no-decode-eq --no-cpu-baseline > gpurun_out/c10_bench_olddgrad.json 2> gpurun_out/c10_bench_olddgrad.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/r2_launches_train.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c10_ncu_list.log 2>&1
tail -6 gpurun_out/c10_pytest.log
python - <<'PY'
import json
for f in ("c10_bench", "c10_bench_olddgrad"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], "train", d["train_step"]["ms_per_step"], d["train_step"]["stages_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
