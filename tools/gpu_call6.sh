#!/bin/bash
# round-2 GPU call 6: validation of the backward tweaks (conv1 wgrad FFMA2, pipelined bn_bwd_reduce, pooled bias sums) + launch list of a training step
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 400 python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
echo "bench rc=$?" >> gpurun_out/c6_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 400 --csv --log-file gpurun_out/r2_launches_train.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c6_ncu_list.log 2>&1
python __graft_entry__.py > gpurun_out/c6_build.log 2>&1; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.log 2>&1
tail -6 gpurun_out/c6_pytest.log; tail -2 gpurun_out/c6_smoke.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c6_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["variants"])
print("train", d["train_step"]["ms_per_step"], d["train_step"]["stages_ms"])
print(d["train_step"].get("forward_stages_train_mode_ms"))
print(d["stages"])
PY
