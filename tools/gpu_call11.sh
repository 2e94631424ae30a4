#!/bin/bash
# round-2 GPU call 11: ncu --set full on one training step's weight-gradient GEMMs, BPTT and the swapped conv2 dgrad
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|lstm_bwd|conv2_dgrad" -s 39 -c 13 -o gpurun_out/r2_full_bwd_gemms \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c11_ncu.log 2>&1
tail -3 gpurun_out/c11_ncu.log | cut -c1-200
