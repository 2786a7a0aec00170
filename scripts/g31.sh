cd /root/repo
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_N$n.json 2> gpurun_out/bench_N$n.err; echo "rc=$?"; tail -2 gpurun_out/bench_N$n.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_N$n.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'], d['n_gpus'], d['scaling'])
PY
done
