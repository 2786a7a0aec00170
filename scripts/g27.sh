cd /root/repo
for cfg in "BN_HEAD_CTAS=3" "BN_HEAD_CTAS=2"; do
  echo "== $cfg"
  env $cfg BN_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bq.json 2> gpurun_out/bq.err; grep "prep_ds: C=" gpurun_out/bq.err | sort | uniq | head -12
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'])
print(' '.join(f"{k.split('_')[0]}_{k.split('_')[2] if k.startswith('K45') else ''}={v}" for k,v in d['roofline']['kernels_ms'].items()))
PY
done
BN_HEAD_CTAS=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
