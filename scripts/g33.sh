cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_evaluate.py -m gpu -x -q 2>&1 | tail -2
for w in 0 9472 21760; do
  echo "== wave $w (0 = engine default)"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --wave $w > gpurun_out/bq_$w.json 2> gpurun_out/bq.err; tail -1 gpurun_out/bq.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bq_$w.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'], 'wave', d['config']['wave'])
PY
done
