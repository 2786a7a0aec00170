cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_N1.json 2> gpurun_out/bench_N1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_N1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_N1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
print(json.dumps(d['roofline']))
PY
