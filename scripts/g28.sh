cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_N1.json 2> gpurun_out/bench_N1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_N1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_N1.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_run_wave_2368x3.csv python profiles/run_wave.py 2368 3 > gpurun_out/rw.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -c 20 -o gpurun_out/full_wave2368 python profiles/run_wave.py 2368 1 > gpurun_out/rw_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench_frontend.py --out gpurun_out/frontend_sweep_config2.json > gpurun_out/fs.log 2>&1; echo "frontend sweep rc=$?"; tail -3 gpurun_out/fs.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_N1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
print(json.dumps(d['roofline']))
print(json.dumps(d.get('cpu_baseline')), d.get('gpu_launches'), d.get('clocks'))
PY
