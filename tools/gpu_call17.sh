#!/bin/bash
# round-2 GPU call 17: full GPU suite on the current build; bench (e2e with the integer feeds pre-staged too); host profile; ncu --set full
# of the tensor-core conv1 weight-gradient kernel and the two data-gradient GEMMs with fused ReLU / BN-sum epilogues
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1300 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c17_pytest.log
tail -6 gpurun_out/c17_pytest.log
timeout 500 python bench.py > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
echo "bench rc=$?" >> gpurun_out/c17_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c17_bench.json") if l.startswith("{")][-1])
    t = d.get("train_step") or {}
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k: v["value"] for k, v in d["e2e"]["variants"].items()}, "train", t.get("ms_per_step"),
          "delta", d["ctc_loss_delta"]["rel"], "decode", d["decode_equality"].get("agreement_unfiltered"))
    print(t.get("stages_ms"))
except Exception as e:
    print("bench failed", e)
PY
timeout 300 python tools/e2e_profile.py 30 > gpurun_out/c17_e2e_profile.txt 2>&1
head -12 gpurun_out/c17_e2e_profile.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"conv1_wgrad_tc|gemm2_kernelILi256ELi1ELi1[34]E" -c 3 -o gpurun_out/r2_full_bwd_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c17_ncu.log 2>&1
tail -2 gpurun_out/c17_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
