cd /root/repo
for cfg in "4736 4736" "3552 3552" "2368 2368"; do
  set -- $cfg
  echo "== subwave $1 wave $2"
  env BN_FE_SUBWAVE=$1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --wave $2 > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -2 gpurun_out/bq.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e']['value'])
print(' '.join(f"{k.split('_')[0]}_{k.split('_')[2] if k.startswith('K45') else ''}={v}" for k,v in d['roofline']['kernels_ms'].items()))
PY
done
