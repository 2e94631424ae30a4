#!/bin/bash
# round-2 GPU call 7: coalesced LSTM saved-state layout; learning test; bench; ncu of the weight-gradient GEMMs + BPTT
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
timeout 400 python bench.py > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
echo "bench rc=$?" >> gpurun_out/c7_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|lstm_bwd|lstm_mc" -s 12 -c 14 -o gpurun_out/r2_full_wgrad_bptt \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c7_ncu.log 2>&1
tail -6 gpurun_out/c7_pytest.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c7_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["variants"])
print("train", d["train_step"]["ms_per_step"], d["train_step"]["stages_ms"])
print(d["train_step"].get("forward_stages_train_mode_ms"))
PY
