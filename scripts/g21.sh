cd /root/repo
for hw in 148 296 592 1184 2368; do
  echo "== host wave $hw"
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --host-wave $hw > gpurun_out/bq.json 2> gpurun_out/bq.err; tail -2 gpurun_out/bq.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print('value',d['value'],'ms',d['ms_per_step'], 'e2e', d['e2e'])
PY
done
