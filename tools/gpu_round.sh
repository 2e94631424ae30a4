#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms), ncu launch list, ncu full capture of the hot kernels.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full (one launch each: conv1_tc, conv2_swap, conv4_1/conv4_2, lstm_mc, ctc_fast)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"conv1_tc|conv2_swap|gemm2_kernelILi256ELi1ELi5E|lstm_mc|ctc_fast" -s 18 -c 6 -o gpurun_out/prof_r1b -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out | tail -12
