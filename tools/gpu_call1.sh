#!/bin/bash
# round-2 GPU call 1: full GPU test suite (with the new benchmark-shape parity cases), training demo on fresh renders, bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
(nvidia-smi -L; nproc; free -g | head -2) > gpurun_out/c1_box.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 420 python tools/train_demo.py --iters 40000 --lr 1e-4 --batch 64 --seconds 270 --tag ref_cfg --save > gpurun_out/c1_train_ref.log 2>&1
echo "train rc=$?" >> gpurun_out/c1_train_ref.log
timeout 300 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?" >> gpurun_out/c1_bench.err
timeout 300 python tools/train_demo.py --iters 40000 --lr 5e-4 --batch 64 --seconds 150 --tag lr5e-4 > gpurun_out/c1_train_lr5.log 2>&1
echo "train rc=$?" >> gpurun_out/c1_train_lr5.log
tail -5 gpurun_out/c1_pytest.log; tail -3 gpurun_out/c1_train_ref.log; tail -c 600 gpurun_out/c1_bench.json; tail -3 gpurun_out/c1_train_lr5.log
