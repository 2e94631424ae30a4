mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), {k:v['ms'] for k,v in d['stages'].items()})
PY
}
echo "== conv1 tc tests"; CRNN_CONV1=tc timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -q -k "forward_layers or golden or gradients_vs_oracle" 2>&1 | tail -3
echo "== bench conv1 tc + lstm ms"; CRNN_LSTM_IMPL=ms CRNN_CONV1=tc timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_tc2.json 2>gpurun_out/bench_tc2.err; show gpurun_out/bench_tc2.json
echo "== e2e probe"; timeout 200 python tools/e2e_probe.py 2>&1 | tail -9
