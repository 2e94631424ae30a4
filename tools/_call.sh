mkdir -p gpurun_out
q() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), {k:v['ms'] for k,v in d['stages'].items()})
PY
}
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for impl in ms ms16; do
  echo "== quick bench lstm=$impl"; CRNN_LSTM_IMPL=$impl timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_q_$impl.json 2>gpurun_out/bench_q_$impl.err; q gpurun_out/bench_q_$impl.json
done
W=$(python - <<'PY'
import json
a=json.load(open('gpurun_out/bench_q_ms.json'))['stages']['lstm_recurrence']['ms']
b=json.load(open('gpurun_out/bench_q_ms16.json'))['stages']['lstm_recurrence']['ms']
print('ms16' if b < a else 'ms')
PY
)
echo "== winner $W: full bench"; CRNN_LSTM_IMPL=$W timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; q gpurun_out/bench.json; echo "winner=$W" > gpurun_out/lstm_winner.txt
echo "== train-mode lstm check ($W)"; CRNN_LSTM_IMPL=$W timeout 200 python -m pytest tests/test_gpu_training.py -q -k "gradients_vs_oracle or three_training" 2>&1 | tail -2
