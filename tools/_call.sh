mkdir -p gpurun_out
echo "== training tests with K-split BPTT"
CRNN_BPTT=ks timeout 400 python -m pytest tests/test_gpu_training.py -q 2>&1 | tail -5
for v in ring ks; do
  echo "== bench train step bptt=$v"
  CRNN_BPTT=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bptt_$v.json 2>gpurun_out/bench_bptt_$v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/bench_bptt_{sys.argv[1]}.json'))
t=d['train_step']; print(round(d['value']), t['ms_per_step'], round(t['images_per_s']), {k:v for k,v in t['stages_ms'].items() if 'lstm' in k or 'forward' in k})
PY
done
tail -2 gpurun_out/bench_bptt_ks.err
