mkdir -p gpurun_out
q() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'lstm', d['stages']['lstm_recurrence']['ms'])
PY
}
echo "== gx tests"
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "cluster_lstm" 2>&1 | tail -3
CRNN_LSTM_IMPL=gx timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -q -k "golden or full_size_c3 or (gradients_vs_oracle and ks) or three_training" 2>&1 | tail -3
for impl in ms gx; do
  echo "== quick bench lstm=$impl"; CRNN_LSTM_IMPL=$impl timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_q_$impl.json 2>gpurun_out/bench_q_$impl.err; q gpurun_out/bench_q_$impl.json
done
echo "== trace gx"; CRNN_LSTM_IMPL=gx timeout 100 python tools/lstm_trace.py 2>&1 | grep lstm_trace | head -3
