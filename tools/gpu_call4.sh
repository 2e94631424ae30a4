#!/bin/bash
# round-2 GPU call 4: TMA-tile CTC kernel + pooled masked bias sums: full suite, bench, ncu of the CTC kernels, backward launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 400 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench rc=$?" >> gpurun_out/c4_bench.err
timeout 200 python tools/ctc_bench.py > gpurun_out/c4_ctc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_ -s 4 -c 2 -o gpurun_out/r2_full_ctc python tools/ctc_bench.py > gpurun_out/c4_ncu_ctc.log 2>&1
# backward launch list + full set on the HBM-bound backward kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd|unpool|colsum|relu_bwd|conv1_wgrad|clip_adam|grad_finish|dlogits" -s 40 -c 16 -o gpurun_out/r2_full_bwd_elem \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c4_ncu_bwd.log 2>&1
tail -6 gpurun_out/c4_pytest.log; tail -c 900 gpurun_out/c4_bench.json; echo; cat gpurun_out/c4_ctc.log
