#!/bin/bash
# round-2 GPU call 13: kind::tf32 configuration (compute_dtype 3) -- parity vs the fp64 oracle, 10k-line decode, bench next to the split path
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_gpu_x3.py tests/test_gpu_decode10k.py -m gpu -q --timeout=500 > gpurun_out/c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
tail -25 gpurun_out/c13_pytest.log
timeout 400 python bench.py --workload c2tf32 > gpurun_out/c13_bench_c2tf32.json 2> gpurun_out/c13_bench_c2tf32.err
echo "bench rc=$?" >> gpurun_out/c13_bench_c2tf32.err
timeout 400 python bench.py --workload c2 > gpurun_out/c13_bench_c2.json 2> gpurun_out/c13_bench_c2.err
echo "bench rc=$?" >> gpurun_out/c13_bench_c2.err
python - <<'PY'
import json
for f in ("c13_bench_c2tf32", "c13_bench_c2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["dtype"], "delta", d.get("ctc_loss_delta", {}).get("rel"), d.get("ctc_loss_delta", {}).get("max_logit_err_rel"),
              "decode", d.get("decode_equality", {}).get("agreement_unfiltered"), "e2e", d["e2e"]["value"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
grep -h "tf32" gpurun_out/parity_report.jsonl | head -20
