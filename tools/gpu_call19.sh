#!/bin/bash
# round-2 GPU call 19: final build -- full GPU suite, default bench line (what the driver runs), smoke(), c2tf32 bench, ncu --set full of the
# kind::tf32 implicit-GEMM kernel
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1300 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c19_pytest.log
tail -5 gpurun_out/c19_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c19_smoke.log 2>&1; tail -2 gpurun_out/c19_smoke.log
timeout 500 python bench.py > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
echo "bench rc=$?" >> gpurun_out/c19_bench.err
timeout 400 python bench.py --workload c2tf32 --no-cpu-baseline --no-decode-eq > gpurun_out/c19_bench_c2tf32.json 2> gpurun_out/c19_bench_c2tf32.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c19_bench.json") if l.startswith("{")][-1])
    t = d.get("train_step") or {}
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k: v["value"] for k, v in d["e2e"]["variants"].items()}, "train", t.get("ms_per_step"),
          "delta", d["ctc_loss_delta"]["rel"], "decode", d["decode_equality"].get("agreement_unfiltered"), "cpu", d["cpu_baseline"]["value"], d["clocks"])
    d = json.loads([l for l in open("gpurun_out/c19_bench_c2tf32.json") if l.startswith("{")][-1])
    print("c2tf32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"gemm_kernelILi256ELi1ELi12ELi4ELi1E" -s 4 -c 2 -o gpurun_out/r2_full_tf32_conv \
    python bench.py --workload c2tf32 --steps 2 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c19_ncu.log 2>&1
tail -2 gpurun_out/c19_ncu.log | cut -c1-200
