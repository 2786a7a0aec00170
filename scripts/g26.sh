cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_features.py tests/test_ingest.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -2 gpurun_out/bq.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'])
print(' '.join(f"{k.split('_')[0]}_{k.split('_')[2] if k.startswith('K45') else ''}={v}" for k,v in d['roofline']['kernels_ms'].items()))
PY
