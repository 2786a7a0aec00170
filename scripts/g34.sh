cd /root/repo
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_N1.json 2> gpurun_out/bench_N1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_N1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_N1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'], d['config']['wave'], d['e2e']['matches_device_run'])
print(json.dumps(d['roofline']))
PY
