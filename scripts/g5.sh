set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r1_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"; cat gpurun_out/r1_bench.json; tail -3 gpurun_out/r1_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2>&1; cat gpurun_out/r1_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/r1_launches.csv python profiles/run_wave.py 2048 3 > gpurun_out/r1_ncu_launch.log 2>&1; tail -2 gpurun_out/r1_ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 37 -c 37 -o /tmp/r1_full python profiles/run_wave.py 2048 2 > gpurun_out/r1_ncu_full.log 2>&1; tail -2 gpurun_out/r1_ncu_full.log
ncu -i /tmp/r1_full.ncu-rep --page raw --csv > gpurun_out/r1_full_raw.csv
ls -la gpurun_out
