#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list, ncu full capture of the top kernels.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== diag (fast fail)"; timeout 180 python tools/diag.py > gpurun_out/diag.log 2>&1 || { echo "DIAG FAILED rc=$?"; tail -30 gpurun_out/diag.log; exit 1; }; tail -28 gpurun_out/diag.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"gemm2_kernelILi1ELi5E|ctc_loss|lstm_persistent|gemm_kernelILi128ELi1ELi3E" -s 15 -c 5 -o gpurun_out/prof_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
