cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_N2.json 2> gpurun_out/bench_N2.err; echo "rc=$?"; tail -2 gpurun_out/bench_N2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_N2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'], d['n_gpus'], d['scaling'])
PY
