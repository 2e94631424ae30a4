#!/bin/bash
# round-2 GPU call 21: pageable-batch path (crnn_forward_pageable: pooled host copy into pinned staging, pipelined with the DMA and the
# conv front end) -- tests + the bench line's "fresh pageable array every step" variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -k "pageable or session or Session or smoke or forward_layers" > gpurun_out/c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c21_pytest.log
tail -5 gpurun_out/c21_pytest.log
for t in 8 16; do
CRNN_HOST_COPY_THREADS=$t timeout 400 python bench.py --no-decode-eq --no-cpu-baseline --no-train > gpurun_out/c21_bench_t$t.json 2> gpurun_out/c21_bench_t$t.err
done
CRNN_H2D_CHUNKS=8 CRNN_HOST_COPY_THREADS=8 timeout 400 python bench.py --no-decode-eq --no-cpu-baseline --no-train > gpurun_out/c21_bench_c8.json 2> gpurun_out/c21_bench_c8.err
python - <<'PY'
import json
for f in ("t8", "t16", "c8"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c21_bench_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k: (v["value"], v.get("path")) for k, v in d["e2e"]["variants"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -2 gpurun_out/c21_bench_t8.err
