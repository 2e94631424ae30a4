mkdir -p gpurun_out
echo "== full suite (defaults: conv2 swap, conv1 FFMA2, STG.256 epilogues, CTC fast, LSTM v1)"
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -12
for impl in mc ds; do
  echo "== LSTM impl $impl: model-level tests"
  CRNN_LSTM_IMPL=$impl timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -q -k "golden or gradients_vs_oracle or full_size_c3 or three_training" 2>&1 | tail -4
  echo "== LSTM impl $impl with LBO/SBO swapped (informative)"
  CRNN_LSTM_SWAPLS=1 timeout 100 python -m pytest tests/test_gpu_parity.py -q -k "cluster_lstm_kernels and $impl" 2>&1 | tail -2
  echo "== trace $impl"; CRNN_LSTM_IMPL=$impl timeout 100 python tools/lstm_trace.py 2>&1 | grep lstm_trace | head -4
  echo "== bench $impl"; CRNN_LSTM_IMPL=$impl timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_$impl.json 2>gpurun_out/bench_$impl.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$impl.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})
PY
done
echo "== bench v1"; timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_v1.json 2>gpurun_out/bench_v1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v1.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})
PY
