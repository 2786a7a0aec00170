cd /root/repo
for cfg in "BN_DS_CTAS=2" "BN_DS_CTAS=3"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -2 gpurun_out/bq.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print('value',d['value'],'ms',d['ms_per_step'])
print(' '.join(f"{k.split('_')[0]}_{k.split('_')[2] if k.startswith('K45') else ''}={v}" for k,v in d['roofline']['kernels_ms'].items()))
PY
done
