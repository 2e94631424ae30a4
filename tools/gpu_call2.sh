#!/bin/bash
# round-2 GPU call 2: full suite (f32-class path, 10k-line decode equality, benchmark-shape parity), both bench workloads, ncu passes
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 400 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?" >> gpurun_out/c2_bench.err
timeout 400 python bench.py --workload c2 > gpurun_out/c2_bench_f32.json 2> gpurun_out/c2_bench_f32.err
echo "bench rc=$?" >> gpurun_out/c2_bench_f32.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err
# launch list of the bench command (device time per launch; shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-decode-eq > gpurun_out/c2_ncu_list.log 2>&1
# full-set capture of one forward step's kernels + the CTC kernel (warm: skip the first ~6 steps)
timeout 900 ncu --set full --clock-control none --import-source on -s 130 -c 22 -o gpurun_out/r2_full_fwd_step \
    python bench.py --steps 3 --warmup 6 --no-cpu-baseline --no-decode-eq --no-train > gpurun_out/c2_ncu_full.log 2>&1
tail -8 gpurun_out/c2_pytest.log; tail -c 1500 gpurun_out/c2_bench.json; echo; tail -c 1200 gpurun_out/c2_bench_f32.json; tail -3 gpurun_out/c2_bench_f32.err
