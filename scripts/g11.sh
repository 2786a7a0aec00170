cd /root/repo
BN_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -v "^prep_ds: C=" | tail -25
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
for k,v in d['roofline']['kernels_ms'].items(): print(k,v)
PY
